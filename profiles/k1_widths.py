"""K1 (k-strongest) at the two real image widths: 3360 bins (MulRan / BASELINE shape, rows 16-byte aligned) and 3768 bins
(Oxford Radar RobotCar: 3768 % 16 = 8, every other row starts 8 bytes off a 16-byte boundary -> the kernel's unaligned
instantiation).  Device-resident images, stage events of the library; prints ms per launch and GB/s of image bytes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfear_radarodometry_code_public_b200 import capi, synth  # noqa: E402

nscan = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for R in (3360, 3768):
    base = np.stack([synth.render_polar(synth.make_world(s), (0.0, 0.0, 0.0), 1000 + s, 400, R) for s in range(8)])
    img = np.ascontiguousarray(np.tile(base, (nscan // 8, 1, 1)))
    ctx = capi.Context(device=0, max_batch=nscan, azimuths=400, range_bins=R, max_cellsets=nscan + 1, max_keyframes=1)
    d_img = ctx.dev_alloc(img.nbytes); ctx.h2d(d_img, img)
    kf = np.full((nscan, 1), nscan, np.int32); cur = np.arange(nscan, dtype=np.int32)
    poses = np.zeros((nscan, 2, 3)); mot = np.zeros((nscan, 3))
    d = {k: ctx.dev_alloc(v.nbytes) for k, v in dict(kf=kf, cur=cur, poses=poses, mot=mot).items()}
    for k, v in dict(kf=kf, cur=cur, poses=poses, mot=mot).items():
        ctx.h2d(d[k], v)
    d_cov = ctx.dev_alloc(nscan * 36 * 8); d_st = ctx.dev_alloc(nscan * capi.STATS_DTYPE.itemsize)
    for rep in range(2):
        ctx.stage_timing(True)
        for _ in range(50):
            ctx.odometry_step_batch_dev(nscan, d_img, d["mot"], d["kf"], 1, d["cur"], d["poses"], d_cov, d_st)
        ctx.sync()
        n, ms = ctx.stage_timing(False)
    k1 = ms[0] / n
    print("R=%d  %d scans: K1 %.4f ms per launch, %.0f GB/s of image bytes, %.3f ns per KB" % (R, nscan, k1, img.nbytes / k1 / 1e6, k1 * 1e6 / (img.nbytes / 1024)))
    ctx.close()
