"""Time ONE visibility sweep per map size (device-resident buffers, CUDA events, fp64 store) on
both routes: one CTA per pair (grid_sweep 0) and the sweep spread over many CTAs (grid_sweep 2)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import visibility_heuristic_path_planner_b200 as vhp

sizes = [int(a) for a in sys.argv[1:]] or [512, 1000, 2048, 4096, 8192]
dev = torch.device("cuda", 0)
for n in sizes:
    g = np.random.default_rng(n)
    occ = np.ones((1, n, n), np.uint8)
    for _ in range(max(4, int(15 * (n / 1000) ** 2))):
        x, y = int(g.integers(1, n)), int(g.integers(1, n))
        w, h = int(g.integers(n // 10, n // 5 + 1)), int(g.integers(n // 10, n // 5 + 1))
        occ[0, y:y + h, x:x + w] = 0
    free = np.argwhere(occ[0] != 0)
    a = free[len(free) // 3]
    src = torch.tensor([[int(a[1]), int(a[0])]], dtype=torch.int32, device=dev)
    occ_t = torch.from_numpy(occ).to(dev)
    outs, line = [], f"{n:5d}^2:"
    for mode in (0, 2):
        stream = torch.cuda.Stream(dev)
        ctx = vhp.torch_context(0, stream)
        ctx.set_grid_sweep(mode)
        out = torch.empty((1, n, n), dtype=torch.float64, device=dev)
        with torch.cuda.stream(stream):
            ctx.prepare_maps_dev(occ_t) if hasattr(ctx, "prepare_maps_dev") else None
            for _ in range(3):
                ctx.visibility_batch_dev(occ_t, src, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                ctx.visibility_batch_dev(occ_t, src, out)
            e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line += f"  {'cta ' if mode == 0 else 'grid'} {ms:8.3f} ms ({n * n / ms / 1e6:7.1f} Gcells/s)"
        outs.append(out.cpu())
        ctx.close()
    assert torch.equal(outs[0], outs[1])
    print(line, flush=True)
