"""runs the per-frame kernels of ONE rank of a band-sharded run (no exchange) on a single GPU, with band-local inputs
addressed through a virtual full-frame base pointer exactly as bench.py does -- for compute-sanitizer.
usage: band_rank_kernels.py W H world [ranks...]"""
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
for g in ranks:
    ctx = Context(0, stream.cuda_stream)
    bp = BandedPipeline(W, H, g, world, True, ctx, cuda_view(dev), max_disp_rows=max(24, -(-24 * H // 1080)))
    lo, hi = bp.plan.input_rows(g)
    fr = synth.render_frame(W, H, 0, rows=(lo, hi))
    bufs = {"depth": torch.from_numpy(np.ascontiguousarray(fr.depth[lo:hi])).to(dev), "normal": torch.from_numpy(np.ascontiguousarray(fr.normal[lo:hi])).to(dev),
            "albedo": torch.from_numpy(np.ascontiguousarray(fr.albedo[lo:hi])).to(dev), "illum": torch.from_numpy(np.ascontiguousarray(fr.illumination[lo:hi])).to(dev)}
    pitch = {"depth": 4 * W, "normal": 8 * W, "albedo": 4 * W, "illum": 16 * W}
    bp.pipe.bind_inputs(*[bufs[k].data_ptr() - lo * pitch[k] for k in ("depth", "normal", "albedo", "illum")])
    p, plan = bp.pipe, bp.plan
    for f in range(16):
        p.set_frame_constants(f, fr.camera)
        p.accumulator.set_row_range(*plan.accumulate_rows(g, f))
        bp._acc_cmd(p.commands)
        bp._bmfr_cmd(p.commands)
        p.taa.set_row_range(*plan.owned_rows(g, f))
        bp._taa_cmd(p.commands)
        bp._back_cmd(p.commands)
        p.end_frame(fr.camera)
        torch.cuda.synchronize()
    print(f"rank {g}/{world}: input rows [{lo},{hi}) block rows {plan.block_rows(g)} ok", flush=True)
    del bp, ctx
