"""Band sharding: integer geometry on the CPU, and a 2-rank gloo run of the
whole banded chain (through the tests/hostsim emulator) that must reproduce the single-rank result
bit for bit -- i.e. the halo plan delivers every history row a rank reads."""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import backend_params

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("W,H,N", [(1920, 1080, 2), (1920, 1080, 4), (3840, 2160, 8), (1920, 8640, 8), (7680, 4320, 8), (256, 256, 2)])
def test_band_plan_partitions_the_frame(W, H, N):
    from vulkanpbrt_b200.multigpu import BandPlan, block_offset
    plan = BandPlan(W, H, N, taa=True)
    assert plan.block_rows(0)[0] == 0 and plan.block_rows(N - 1)[1] == H // 32 + 2
    for f in range(34):
        oy = block_offset(32, f)[1]
        rows = [plan.owned_rows(g, f) for g in range(N)]
        assert rows[0][0] == 0 and rows[-1][1] == H
        for g in range(N - 1):
            assert rows[g][1] == rows[g + 1][0]                      # contiguous partition of [0, H)
        for g in range(N):
            b0, b1 = plan.block_rows(g)
            # rows written by the rank's BMFR blocks (non-mirrored pixels, bmfrPost.comp:74) lie in its owned rows
            wlo, whi = max(0, 32 * b0 - oy), min(H, 32 * b1 - oy)
            assert rows[g][0] <= wlo and whi <= rows[g][1]
            a = plan.accumulate_rows(g, f)
            assert a[0] <= rows[g][0] and rows[g][1] <= a[1]
            # every image row a block of the rank reads through jitter + mirror is accumulated locally
            for ay in (32 * b0 - oy, 32 * b1 - oy - 1):
                m = -ay - 1 if ay < 0 else (2 * H - ay - 1 if ay >= H else ay)
                assert a[0] <= m < a[1]


def test_everything_sent_after_bmfr_comes_from_the_edge_block_rows():
    """exchange B starts after the edge block rows only: every row it sends must have been written by them"""
    from vulkanpbrt_b200.multigpu import BandPlan, block_offset
    for (W, H, N) in [(1920, 1080, 2), (1920, 2160, 2), (3840, 2160, 8), (1920, 8640, 8)]:
        plan = BandPlan(W, H, N, taa=True)
        ne = plan.edge_block_rows
        for f in range(34):
            oy = block_offset(32, f)[1]
            ts = [t for t in plan.history_transfers(f + 1) if t.plane == "denoised"] + plan.stale_column_transfers(f) + plan.final_transfers(f)
            for t in ts:
                b0, b1 = plan.block_rows(t.src)
                nt, nb = (ne if t.src > 0 else 1), (ne if t.src < N - 1 else plan.bottom_wrap_block_rows)   # as BandedPipeline.run_frame
                top = (max(0, 32 * b0 - oy), min(H, 32 * (b0 + nt) - oy))
                bot = (max(0, 32 * (b1 - nb) - oy), min(H, 32 * b1 - oy))
                # rows above -oy are written by nobody this frame (frame 8: rows 0..1 keep their old content)
                r0, r1 = max(t.rows[0], -oy, 0), t.rows[1]
                if r1 <= r0:
                    continue
                in_top = top[0] <= r0 and r1 <= top[1]
                in_bot = bot[0] <= r0 and r1 <= bot[1]
                assert in_top or in_bot, (W, H, N, f, t, top, bot)


def test_history_transfers_cover_every_needed_row():
    from vulkanpbrt_b200.multigpu import BandPlan
    for (W, H, N) in [(1920, 1080, 4), (640, 2160, 8), (256, 256, 2)]:
        plan = BandPlan(W, H, N, taa=True)
        for f in range(1, 34):
            ts = plan.history_transfers(f)
            for dst in range(N):
                for plane, have, need in (("acc", plan.accumulate_rows(dst, f - 1), plan.accumulate_rows(dst, f)),
                                          ("denoised", plan.owned_rows(dst, f - 1), plan.owned_rows(dst, f)),
                                          ("taa", plan.owned_rows(dst, f - 1), plan.owned_rows(dst, f))):
                    got = np.zeros(H, bool)
                    got[have[0]:have[1]] = True
                    for t in ts:
                        if t.dst == dst and t.plane == plane:
                            o = plan.owned_rows(t.src, f - 1)
                            assert o[0] <= t.rows[0] and t.rows[1] <= o[1]          # sent by the canonical owner
                            got[t.rows[0]:t.rows[1]] = True
                    lo, hi = max(0, need[0] - plan.D - 1), min(H, need[1] + plan.D + 1)
                    assert got[lo:hi].all()
                    if lo == 0:
                        assert got[H - 1]                                          # REPEAT wrap rows
                    if hi == H:
                        assert got[0]


def _worker(rank, world, W, H, frames, use_taa, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from vulkanpbrt_b200 import Context, _capi, synth
    from vulkanpbrt_b200.multigpu import BandedPipeline
    _capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so")))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def view(img):     # emulator "device" memory is host memory
        bv = img.byte_view()
        n = int(np.prod(bv.shape))
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(bv.ptr)).reshape(bv.shape))

    bp = BandedPipeline(W, H, rank, world, use_taa, Context(0), view, max_disp_rows=12, external_inputs=False, dist=dist)
    lo, hi = bp.plan.input_rows(rank)
    finals, denoised = [], []
    for f in range(frames):
        fr = synth.render_frame(W, H, f, rows=(lo, hi))      # the rank only ever sees its band + apron of the inputs
        bp.pipe.upload_frame(fr)
        bp.run_frame(f, fr.camera)
        o = bp.owned_rows(f)
        finals.append((o, bp.pipe.final.download()[o[0]:o[1]].copy()))
        denoised.append((o, bp.bmfr.denoised.download()[(f & 1) ^ 1, o[0]:o[1]].copy()))
    bp.flush()
    np.save(os.path.join(out_dir, f"final_{rank}.npy"), np.array([(o, a) for o, a in finals], dtype=object), allow_pickle=True)
    np.save(os.path.join(out_dir, f"den_{rank}.npy"), np.array([(o, a) for o, a in denoised], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_taa,W,H,frames", [(False, 128, 192, 10), (True, 128, 192, 10), (True, 64, 1080, 27)])
def test_two_rank_banded_chain_equals_single_rank(tmp_path, use_taa, W, H, frames):
    """the 1080-row case moves the band boundary over rows whose column 0 is left unwritten at frames 9 and 25
    (negative x jitter): the stale pixels must come from whichever rank wrote them last"""
    import subprocess
    import torch.multiprocessing as mp
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, W, H, frames, use_taa, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    # single-rank run of the same sequence (emulator, this process)
    from vulkanpbrt_b200 import DenoisePipeline, _capi, synth
    saved = _capi._lib
    _capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so")))
    try:
        pipe = DenoisePipeline(W, H, use_taa=use_taa)
        per_rank_f = [np.load(tmp_path / f"final_{r}.npy", allow_pickle=True) for r in range(world)]
        per_rank_d = [np.load(tmp_path / f"den_{r}.npy", allow_pickle=True) for r in range(world)]
        for f in range(frames):
            pipe.run_frame(f, synth.render_frame(W, H, f))
            full_final = pipe.final.download()
            full_den = pipe.modules[0].denoised.download()[(f & 1) ^ 1]
            for r in range(world):
                (lo, hi), band = per_rank_f[r][f]
                np.testing.assert_array_equal(band, full_final[lo:hi], err_msg=f"final, frame {f}, rank {r}")
                (lo, hi), band = per_rank_d[r][f]
                np.testing.assert_array_equal(band, full_den[lo:hi], err_msg=f"denoised, frame {f}, rank {r}")
        del pipe
    finally:
        import gc
        gc.collect()
        _capi._lib = saved


# ---- real GPUs: halo exchange over NVLink (peer memory / NCCL) ---------------------------------------
def _gpu_worker(rank, world, W, H, frames, use_taa, port, out_dir, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    import time

    import torch
    import torch.distributed as dist
    from vulkanpbrt_b200 import Context, synth
    from vulkanpbrt_b200.multigpu import BandedPipeline, NcclDirect, PeerDirect, cuda_view
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = Context(rank, stream.cuda_stream)
    view = cuda_view(dev)
    bp = BandedPipeline(W, H, rank, world, use_taa, ctx, view, external_inputs=(mode == "peer-async"), dist=dist,
                        nccl=NcclDirect(dist, rank, world, dev) if mode == "nccl" else None,
                        peer=PeerDirect(dist, rank, world, dev, ctx) if mode.startswith("peer") else None)
    lo, hi = bp.plan.input_rows(rank)
    finals = []
    if mode == "peer-async":
        # no host synchronisation between frames and the ranks' hosts deliberately out of step: the only
        # ordering left is the flag protocol.  Owned rows are snapshotted on the stream and read at the end.
        seq = [synth.render_frame(W, H, f, rows=(lo, hi)) for f in range(frames)]
        dseq = [{k: torch.from_numpy(np.ascontiguousarray(getattr(fr, k))).to(dev) for k in ("depth", "normal", "albedo", "illumination")}
                for fr in seq]
        torch.cuda.synchronize()
        snaps = []
        for f, fr in enumerate(seq):
            d = dseq[f]
            bp.pipe.bind_inputs(d["depth"].data_ptr(), d["normal"].data_ptr(), d["albedo"].data_ptr(), d["illumination"].data_ptr())
            bp.run_frame(f, fr.camera)
            o = bp.owned_rows(f)
            layer = (f & 1) ^ 1
            snaps.append((o, view(bp.pipe.final)[o[0]:o[1]].clone(), view(bp.bmfr.denoised)[layer, o[0]:o[1]].clone()))
            if (f + rank) % 3 == 0:
                time.sleep(0.02 * (1 + (rank + f) % 3))
        bp.flush()
        torch.cuda.synchronize()
        bp.check()
        for o, fin, den in snaps:
            finals.append((o, fin.cpu().numpy().view(np.uint8).reshape(o[1] - o[0], W, 4), den.cpu().numpy().view(np.uint16).reshape(o[1] - o[0], W, 4)))
    else:
        for f in range(frames):
            fr = synth.render_frame(W, H, f, rows=(lo, hi))
            bp.pipe.upload_frame(fr)
            bp.run_frame(f, fr.camera)
            torch.cuda.synchronize()
            o = bp.owned_rows(f)
            finals.append((o, bp.pipe.final.download()[o[0]:o[1]].copy(), bp.bmfr.denoised.download()[(f & 1) ^ 1, o[0]:o[1]].copy()))
        bp.flush()
        torch.cuda.synchronize()
        bp.check()
    np.save(os.path.join(out_dir, f"gpu_{rank}.npy"), np.array(finals, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,mode", [(2, "peer"), (2, "peer-async"), (2, "nccl"), (2, "torch"), (4, "peer-async"), (8, "peer-async")])
def test_banded_chain_on_gpus_equals_single_gpu(tmp_path, world, mode):
    """peer: rows stored straight into the neighbours' HBM over NVLink + flag words (the bench path);
    peer-async: the same without any host synchronisation between frames and with the ranks' hosts out of step;
    nccl: NCCL groups issued through ctypes on a communication stream; torch: torch.distributed's own P2P ops"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    W, H, frames = 1920, 1080, (20 if mode == "peer-async" else 12)
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_gpu_worker, args=(world, W, H, frames, True, port, str(tmp_path), mode), nprocs=world, join=True, start_method="spawn")
    from vulkanpbrt_b200 import DenoisePipeline, synth
    pipe = DenoisePipeline(W, H, use_taa=True)
    per_rank = [np.load(tmp_path / f"gpu_{r}.npy", allow_pickle=True) for r in range(world)]
    for f in range(frames):
        pipe.run_frame(f, synth.render_frame(W, H, f))
        full_final = pipe.final.download()
        full_den = pipe.modules[0].denoised.download()[(f & 1) ^ 1]
        for r in range(world):
            (lo, hi), band_f, band_d = per_rank[r][f]
            np.testing.assert_array_equal(band_f, full_final[lo:hi], err_msg=f"final, frame {f}, rank {r}")
            np.testing.assert_array_equal(band_d, full_den[lo:hi], err_msg=f"denoised, frame {f}, rank {r}")


# ---- the NVLink peer-memory exchange protocol on the CPU: ranks are THREADS over the test emulator ----------------
class _ThreadGroup:
    """all_gather_object for ranks that are threads of this process"""

    def __init__(self, world):
        import threading
        self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)

    def rank_view(self, rank):
        group = self

        class _Dist:
            def all_gather_object(self, out, obj):
                group.slots[rank] = obj
                group.barrier.wait()
                out[:] = list(group.slots)
                group.barrier.wait()
        return _Dist()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,use_taa,W,H,frames", [(2, True, 64, 416, 20), (3, True, 64, 416, 20), (3, False, 96, 384, 12),
                                                       (2, True, 101, 400, 12)])      # odd width: unaligned rows take the 4-byte / 1-byte copy paths
def test_peer_memory_exchange_protocol_on_the_emulator(world, use_taa, W, H, frames):
    """PeerDirect + vkpbrt_halo_exchange_* + k_halo_push / k_halo_wait, compiled for the CPU: every rank is an OS thread
    with its own context and pipeline, "peer" memory is the shared address space, flag words are real atomics and the
    ranks run at whatever speed the scheduler gives them.  The banded result must equal the single-rank one bit for
    bit (a protocol error shows up as stale halo rows, or as a flag wait running into its timeout)."""
    import subprocess
    import threading
    import time
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    from vulkanpbrt_b200 import Context, DenoisePipeline, _capi, synth
    from vulkanpbrt_b200.multigpu import BandedPipeline, PeerDirect
    saved = _capi._lib
    _capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so")))
    try:
        group = _ThreadGroup(world)
        results, errors = [None] * world, []

        def rank_main(rank):
            try:
                ctx = Context(0)
                peer = PeerDirect(group.rank_view(rank), rank, world, None, ctx, timeout_ms=120000, comm_stream=0)
                bp = BandedPipeline(W, H, rank, world, use_taa, ctx, None, max_disp_rows=12, external_inputs=False, peer=peer)
                lo, hi = bp.plan.input_rows(rank)
                out = []
                for f in range(frames):
                    fr = synth.render_frame(W, H, f, rows=(lo, hi))
                    bp.pipe.upload_frame(fr)
                    bp.run_frame(f, fr.camera)
                    o = bp.owned_rows(f)
                    out.append((o, bp.pipe.final.download()[o[0]:o[1]].copy(), bp.bmfr.denoised.download()[(f & 1) ^ 1, o[0]:o[1]].copy()))
                    if (f + rank) % 3 == 0:         # keep the ranks out of step
                        time.sleep(0.004 * (1 + (rank + f) % 4))
                bp.flush()
                bp.check()
                results[rank] = (out, bp, peer)       # handles stay alive until every rank is done writing into them
            except BaseException as e:      # noqa: BLE001 -- reported by the main thread
                errors.append((rank, e))
                group.barrier.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        pipe = DenoisePipeline(W, H, use_taa=use_taa)
        for f in range(frames):
            pipe.run_frame(f, synth.render_frame(W, H, f))
            full_final = pipe.final.download()
            full_den = pipe.modules[0].denoised.download()[(f & 1) ^ 1]
            for r in range(world):
                (lo, hi), band_f, band_d = results[r][0][f]
                np.testing.assert_array_equal(band_f, full_final[lo:hi], err_msg=f"final, frame {f}, rank {r}")
                np.testing.assert_array_equal(band_d, full_den[lo:hi], err_msg=f"denoised, frame {f}, rank {r}")
        for r in range(world):
            results[r][2].close()
        del pipe, results
    finally:
        import gc
        gc.collect()
        _capi._lib = saved


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,use_taa,W,H,frames", [(2, True, 64, 416, 20), (3, True, 64, 480, 18), (3, False, 96, 384, 10)])
def test_native_banded_rank_on_the_emulator(world, use_taa, W, H, frames):
    """vkpbrt::BandedRank (include/vkpbrt/banded.hpp) through the C ABI (vkpbrt_banded_rank_*): the C++ host of the
    band-sharded chain, one call per frame, band-local inputs bound through virtual full-frame base pointers.  Ranks are
    OS threads over the emulator; the all_gather callback is a thread barrier.  Owned rows must equal the single-rank run
    bit for bit."""
    import subprocess
    import threading
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    from vulkanpbrt_b200 import Context, DenoisePipeline, _capi, synth
    from vulkanpbrt_b200.multigpu import NativeBandedRank
    saved = _capi._lib
    _capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so")))
    try:
        group = _ThreadGroup(world)
        results, errors, keep = [None] * world, [], [None] * world

        def rank_main(rank):
            try:
                ctx = Context(0)
                nr = NativeBandedRank(W, H, rank, world, use_taa, ctx, dist=group.rank_view(rank), max_disp_rows=12, external_inputs=True,
                                      comm_stream=0, timeout_ms=120000)
                lo, hi = nr.input_rows()
                out = []
                for f in range(frames):
                    fr = synth.render_frame(W, H, f, rows=(lo, hi))
                    planes = [np.ascontiguousarray(fr.depth[lo:hi]), np.ascontiguousarray(fr.normal[lo:hi]), np.ascontiguousarray(fr.albedo[lo:hi]),
                              np.ascontiguousarray(fr.illumination[lo:hi])]
                    pitch = [4 * W, 8 * W, 4 * W, 16 * W]
                    nr.bind_inputs(*[p.ctypes.data - lo * pt for p, pt in zip(planes, pitch)])
                    nr.run_frame(f, NativeBandedRank.camera_block(fr.camera))
                    o = nr.owned_rows(f)
                    out.append((o, nr.final.download()[o[0]:o[1]].copy(), nr.denoised.download()[(f & 1) ^ 1, o[0]:o[1]].copy()))
                nr.flush()
                nr.check()
                results[rank], keep[rank] = out, (nr, ctx)     # handles stay alive until every rank is done writing into them
            except BaseException as e:      # noqa: BLE001 -- reported by the main thread
                import traceback
                traceback.print_exc()
                errors.append((rank, e))
                group.barrier.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        pipe = DenoisePipeline(W, H, use_taa=use_taa)
        for f in range(frames):
            pipe.run_frame(f, synth.render_frame(W, H, f))
            full_final = pipe.final.download()
            full_den = pipe.modules[0].denoised.download()[(f & 1) ^ 1]
            for r in range(world):
                (lo, hi), band_f, band_d = results[r][f]
                np.testing.assert_array_equal(band_f, full_final[lo:hi], err_msg=f"final, frame {f}, rank {r}")
                np.testing.assert_array_equal(band_d, full_den[lo:hi], err_msg=f"denoised, frame {f}, rank {r}")
        for k in keep:
            k[0].close()
        del pipe, results, keep
    finally:
        import gc
        gc.collect()
        _capi._lib = saved


def _pitched(camera, angle):
    """the camera turned about its own x axis: the image content moves vertically by ~ angle / fov_y * H rows"""
    import copy
    c = copy.copy(camera)
    V = np.asarray(camera.view, np.float64).reshape(4, 4).T
    R = np.eye(4)
    R[1:3, 1:3] = [[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]]
    V2 = R @ V
    c.view = V2.T.astype(np.float32).reshape(-1).copy()
    c.inv_view = np.linalg.inv(V2).T.astype(np.float32).reshape(-1).copy()
    return c


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_reprojection_beyond_the_halo_is_counted(backend):
    """the guard behind band sharding's one assumption (a rank holds history rows only within max_disp_rows of the rows it
    computes): k_accumulate counts every reprojection tap that lands further away, instead of serving stale rows.  The
    synthetic sequence's own motion stays inside 12 rows; a camera that pitches by 0.3 rad between two frames does not."""
    from vulkanpbrt_b200 import DenoisePipeline, synth
    W, H = 160, 256
    pipe = DenoisePipeline(W, H, use_taa=False)
    pipe.accumulator.set_max_displacement_rows(12)
    for f in range(3):
        pipe.run_frame(f, synth.render_frame(W, H, f))
    pipe.ctx.synchronize()
    assert pipe.accumulator.displacement_violations() == 0
    fr = synth.render_frame(W, H, 3)
    fr.camera = _pitched(fr.camera, 0.3)
    pipe.run_frame(3, fr)
    pipe.ctx.synchronize()
    n = pipe.accumulator.displacement_violations()
    assert n > W * H // 8, n                      # most pixels whose reprojection stays on screen
    pipe.accumulator.set_max_displacement_rows(0)             # off: nothing is counted
    fr = synth.render_frame(W, H, 4)
    fr.camera = _pitched(fr.camera, -0.3)
    pipe.run_frame(4, fr)
    pipe.ctx.synchronize()
    assert pipe.accumulator.displacement_violations() in (0, n)


@pytest.mark.timeout(600)
def test_native_banded_rank_fails_loudly_when_the_camera_outruns_the_halo():
    """vkpbrt_banded_rank_check after a camera jump: every rank reports the violation (VKPBRT error with the count and
    the remedy) instead of returning frames built from stale history rows"""
    import subprocess
    import threading
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    from vulkanpbrt_b200 import Context, VkpbrtError, _capi, synth
    from vulkanpbrt_b200.multigpu import NativeBandedRank
    saved = _capi._lib
    _capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so")))
    W, H, world = 64, 416, 2
    try:
        group = _ThreadGroup(world)
        messages, errors, keep = [None] * world, [], [None] * world

        def rank_main(rank):
            try:
                ctx = Context(0)
                nr = NativeBandedRank(W, H, rank, world, True, ctx, dist=group.rank_view(rank), max_disp_rows=12, external_inputs=True,
                                      comm_stream=0, timeout_ms=120000)
                lo, hi = nr.input_rows()
                for f in range(4):
                    fr = synth.render_frame(W, H, f, rows=(lo, hi))
                    cam = fr.camera if f < 3 else _pitched(fr.camera, 0.3)
                    planes = [np.ascontiguousarray(fr.depth[lo:hi]), np.ascontiguousarray(fr.normal[lo:hi]), np.ascontiguousarray(fr.albedo[lo:hi]),
                              np.ascontiguousarray(fr.illumination[lo:hi])]
                    nr.bind_inputs(*[p.ctypes.data - lo * pt for p, pt in zip(planes, [4 * W, 8 * W, 4 * W, 16 * W])])
                    nr.run_frame(f, NativeBandedRank.camera_block(cam))
                    if f == 2:
                        nr.flush()
                        nr.check()                     # the sequence's own motion is covered
                nr.flush()
                try:
                    nr.check()
                except VkpbrtError as e:
                    messages[rank] = str(e)
                keep[rank] = (nr, ctx)
            except BaseException as e:      # noqa: BLE001 -- reported by the main thread
                errors.append((rank, e))
                group.barrier.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        for m in messages:
            assert m is not None and "max_disp_rows" in m, m
        for k in keep:
            k[0].close()
        del keep
    finally:
        import gc
        gc.collect()
        _capi._lib = saved
