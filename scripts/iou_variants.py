"""Developer tool: A/B of library build variants on the pairwise-IoU workloads.  Times ops.iou_aabb on three inputs for
every mydetection_b200/_tune/libmydet_<tag>.so present (one process per variant, selected through MYDET_LIB) and for the
default library.  The table in profiles/r2_history.md ("Late finding") was made with five temporary variants of
iou_aabb_kernel (columns per thread x how the divide is skipped), built with `build.build(extra_flags=..., out=...)`."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'one':
    sys.path.insert(0, ROOT)
    import torch
    from mydetection_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(0)
    n, k = 16384, 8192

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    dense_a = (torch.rand(n, 4, generator=g) * 200 + 20).to(dev)
    dense_b = (torch.rand(k, 4, generator=g) * 200 + 20).to(dev)
    sp_a = torch.cat([torch.rand(n, 2, generator=g) * 1000, torch.rand(n, 2, generator=g) * 60 + 4], 1).to(dev)
    sp_b = torch.cat([torch.rand(k, 2, generator=g) * 1000, torch.rand(k, 2, generator=g) * 60 + 4], 1).to(dev)
    r = [timed(lambda: ops.iou_aabb(dense_a, dense_b)), timed(lambda: ops.iou_aabb(sp_a, sp_b)),
         timed(lambda: ops.iou_aabb(sp_a[:8525], sp_b[:100]), 50)]
    print(f'{os.environ.get("MYDET_LIB", "default")[-16:]:18s} dense {r[0]:7.1f} us   sparse {r[1]:7.1f} us   8525x100 {r[2]:6.1f} us', flush=True)
else:
    libs = sorted(glob.glob(os.path.join(ROOT, 'mydetection_b200', '_tune', 'libmydet_*.so'))) + [None]
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env['MYDET_LIB'] = lib
        subprocess.run([sys.executable, __file__, 'one'], env=env, check=False)
