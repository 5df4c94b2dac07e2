"""Wall time of single-mesh octree builds (the adapter's path: one mesh, host round trip per level) with and without the top phase."""
import time, sys
import numpy as np
from newmsm_b200 import capi, synth, resampler as R, build
build.build_library()
L = capi.lib()
for lvl in (4, 5, 6, 7):
    xyz, tri = synth.icosphere(lvl)
    m = R.Mesh(xyz, tri)
    for top in (0, -1):
        capi.check(L.msmgpu_set_tuning(b"build_top", top))
        ts = []
        for _ in range(8):
            t0 = time.perf_counter()
            t = R.Octree(m)
            capi.check(L.msmgpu_ctx_sync(m.ctx.h))
            ts.append(time.perf_counter() - t0)
            t.close()
        print(f"ico{lvl} ({len(tri)} triangles) build_top={top}: median {1e3 * sorted(ts)[len(ts) // 2]:.3f} ms, min {1e3 * min(ts):.3f} ms")
