"""per-kernel GPU time of ONE rank of a band-sharded run (no exchange, single GPU): where a rank's share of the frame goes and how
much the bands differ (the synthetic scene's top quarter is sky: every reprojection is accepted there, so the history paths of
k_bmfr_block's epilogue and k_taa always run).  CUDA events around every launch, median over the timed frames.
usage: band_kernel_times.py W H world [ranks...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vulkanpbrt_b200 import synth
from vulkanpbrt_b200.modules import Context
from vulkanpbrt_b200.multigpu import BandedPipeline, cuda_view

W, H, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ranks = [int(a) for a in sys.argv[4:]] or list(range(world))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
R, WARM, K = 6, 12, 36
seq = lambda f: (f % (2 * R - 2)) if (f % (2 * R - 2)) < R else 2 * R - 2 - (f % (2 * R - 2))     # back and forth: no camera jump
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for g in ranks:
    ctx = Context(0, stream.cuda_stream)
    bp = BandedPipeline(W, H, g, world, True, ctx, cuda_view(dev), max_disp_rows=max(24, -(-24 * H // 1080)))
    lo, hi = bp.plan.input_rows(g)
    pitch = {"depth": 4 * W, "normal": 8 * W, "albedo": 4 * W, "illum": 16 * W}
    frames = []
    for i in range(R):
        fr = synth.render_frame(W, H, i, rows=(lo, hi))
        frames.append((fr.camera, {"depth": torch.from_numpy(np.ascontiguousarray(fr.depth[lo:hi])).to(dev),
                                   "normal": torch.from_numpy(np.ascontiguousarray(fr.normal[lo:hi])).to(dev),
                                   "albedo": torch.from_numpy(np.ascontiguousarray(fr.albedo[lo:hi])).to(dev),
                                   "illum": torch.from_numpy(np.ascontiguousarray(fr.illumination[lo:hi])).to(dev)}))
    p, plan = bp.pipe, bp.plan
    names = ["acc", "bmfr", "taa_inner", "taa_edges"]
    ms = {n: [] for n in names}
    for f in range(WARM + K):
        cam, bufs = frames[seq(f)]
        p.bind_inputs(*[bufs[k].data_ptr() - lo * pitch[k] for k in ("depth", "normal", "albedo", "illum")])
        p.set_frame_constants(f, cam)
        p.accumulator.set_row_range(*plan.accumulate_rows(g, f))
        o0, o1 = plan.owned_rows(g, f)
        p.taa.set_row_range(o0, o1)
        i0, i1 = o0 + (1 if g > 0 else 0), o1 - (1 if g < world - 1 else 0)
        steps = [lambda: bp._acc_cmd(p.commands), lambda: bp._bmfr_cmd(p.commands),
                 lambda: p.taa.record_part(p.push_constants, i0, i1, False),
                 lambda: p.taa.record_parts(p.push_constants, o0, i0, i1, o1, True)]
        flush.fill_(f & 255)       # L2 flush between frames, and ~0.3 ms of head start for the host: the launches below queue up behind it
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(steps) + 1)]
        ev[0].record(stream)
        for i, s in enumerate(steps):
            s()
            ev[i + 1].record(stream)
        bp._back_cmd(p.commands)
        p.end_frame(cam)
        torch.cuda.synchronize()
        if f >= WARM:
            for i, n in enumerate(names):
                ms[n].append(ev[i].elapsed_time(ev[i + 1]))
    med = {n: float(np.median(v)) * 1e3 for n, v in ms.items()}
    print(f"rank {g}/{world} block rows {plan.block_rows(g)} owned(f=0) {plan.owned_rows(g, 0)}: "
          + "  ".join(f"{n} {med[n]:.1f} us" for n in names) + f"   sum {sum(med.values()):.1f} us", flush=True)
    del bp, ctx, frames
