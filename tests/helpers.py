"""Shared builders for the parity tests (seeded synthetic inputs; no reference access at run time)."""
import numpy as np

from cfear_radarodometry_code_public_b200 import synth

A, R = 400, 3360


def adversarial_image(seed=0, A=A, R=R):
    """Rows that stress the tie-break / capacity edges of the k-strongest definition."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.integers(0, 50, (A, R), dtype=np.uint8)
    img[0] = 0                                   # nothing above z_min
    img[1] = 255                                 # saturated row: ties everywhere -> largest ranges win
    img[2] = 60                                  # all exactly z_min
    img[3] = 59                                  # all just below
    img[4, ::7] = 200                            # many equal peaks
    img[5, :5] = [61, 62, 63, 64, 65]            # fewer than k candidates, near range (below min-range cut)
    img[6, -12:] = np.arange(100, 112)           # candidates at the very end of the row
    img[7, :12] = np.arange(111, 99, -1)         # candidates at the very start
    img[8] = rng.integers(0, 256, R)             # dense random: ~77% above z_min (candidate list overflow)
    img[9, 100:400] = 128                        # long plateau
    img[10] = (np.arange(R) % 256).astype(np.uint8)
    img[11, 58] = 200; img[11, 59] = 200         # straddles the min-range bin (58 dropped, 59 kept)
    img[12, 3] = 250; img[12, R - 4] = 250; img[12, R - 3] = 250   # peaks guard band
    for a in range(13, 40):
        n = rng.integers(0, 40)
        cols = rng.integers(0, R, n)
        img[a, cols] = rng.integers(55, 70, n)
    img[40:60] = rng.integers(40, 90, (20, R))   # moderately dense
    return img


def scan_images(seed, n_keyframes=4):
    return synth.make_problem_images(seed, n_keyframes)


def oracle_cells(orc, img, radius=3.5, k=12, z_min=60, weight_intensity=True, mot=None):
    idx, cnt = orc.kstrongest(img, z_min, k)
    cl = orc.cloud(img, idx, cnt)
    if mot is not None:
        cl = orc.compensate(cl, mot)
    return cl, orc.surface_points(cl, radius, weight_intensity)
