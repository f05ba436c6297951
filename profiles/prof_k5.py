import sys, numpy as np
sys.path.insert(0, "/root/repo")
from cfear_radarodometry_code_public_b200 import capi, workload
import ctypes as C
capi.LIB_PATH = "/root/repo/profiles/libcfear_prof.so"
nprob, K = 256, 4
b = workload.make_batch(nprob, K, seed0=0)
ctx = capi.Context(device=0, max_batch=nprob, max_cellsets=nprob*(K+1), max_keyframes=K, **workload.CFEAR3)
kf = np.arange(nprob*K, dtype=np.int32).reshape(nprob, K); cur = (nprob*K+np.arange(nprob)).astype(np.int32)
for i in range(K): ctx.scans_to_cells_batch(b["kf_polar"][:, i], None, kf[:, i])
for _ in range(3): out = ctx.odometry_step_batch(b["polar"], b["mot"], kf, cur, b["poses"])
# now the batched register only (all problems in one launch)
slots = np.concatenate([kf, cur[:, None]], 1)
p, cov, st, _ = ctx.register_batch(slots, b["poses"])
c = cov.reshape(nprob, 36)
sincos, loop, red, nev, build, total = c[:, 8], c[:, 9], c[:, 10], c[:, 11], c[:, 13], c[:, 15]
print("per problem (cycles): total %.0f  evals %.1f  sincos/eval %.0f  loop/eval %.0f  reduce/eval %.0f  build total %.0f (%.1f outer)" % (
    total.mean(), nev.mean(), (sincos/nev).mean(), (loop/nev).mean(), (red/nev).mean(), build.mean(), st["outer_iterations"].mean()))
print("build split per problem: nn %.0f  post-nn %.0f  scan %.0f" % (c[:,14].mean(), c[:,16].mean(), c[:,17].mean()))
print("eval total %.0f, build %.0f, rest(scalar etc) %.0f" % ((sincos+loop+red).mean(), build.mean(), (total-sincos-loop-red-build).mean()))
