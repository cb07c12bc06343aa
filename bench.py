#!/usr/bin/env python
"""bench.py -- BMFR denoised MPix/s and ms/frame on the synthetic G-buffer sequence (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one frame through the hot path: accumulate -> fused BMFR (pre + fit + post) -> TAA.
The default workload, for every N, is BASELINE.json configs[3] -- the full chain at 3840x2160, the configuration the
north star's "< 1 ms per frame" is quoted on.  N = 1 denoises whole frames; N > 1 (torchrun, one rank per GPU) is
STRONG scaling: the same 4K frame cut into horizontal block bands, the history halo rows pushed into the neighbours'
HBM over NVLink every frame (vulkanpbrt_b200/multigpu.py).  The other BASELINE configurations are measured in the same
run and nested under "also": N = 1 adds BMFR 1080p (configs[1]), BFR x3 + blender 1080p (configs[2]) and BMFR 8K;
N = 8 adds the 8K frame band-sharded over the 8 GPUs (configs[4]).

value   whole-job MPix/s with the input sequence already resident in HBM (each frame's planes are
        distinct buffers, read once: inputs larger than L2)
e2e     the same metric through the public API with HOST buffers: per frame the G-buffer + raw
        illumination are copied from pinned host memory and the BGRA8 result is read back, inside the
        timed region (2-deep copy/compute overlap)
roofline  the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
cpu_baseline  the reference's shader source compiled for the host cores (oracle/_ref), bounded sample
--impl reference  times the reference's CPU path (oracle/_ref when built, else the oracle port) alone, on the SAME
        workload at its real size, with every host core
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (W, H, taa, description)
    "bmfr_1080p": (1920, 1080, False, "BMFR 1920x1080 1-spp synthetic sequence with camera motion (BASELINE configs[1])"),
    "bmfr_taa_4k": (3840, 2160, True, "accumulator + BMFR + TAA 3840x2160 1-spp (BASELINE configs[3])"),
    "bmfr_8k": (7680, 4320, False, "BMFR 7680x4320 (BASELINE configs[4])"),
    "bmfr_256": (256, 256, False, "BMFR 256x256 (BASELINE configs[0], plumbing)"),
    "bfr_blend_1080p": (1920, 1080, False, "BFR b=8/16/32 + BFRBlender 1920x1080 1-spp (BASELINE configs[2])"),
}
# algorithmic (compulsory) HBM bytes per image pixel, reference storage formats (DESIGN.md / SURVEY.md 8d)
BYTES_ACCUMULATE = 33 + 17     # reads depth 4 + raw 16 + prev_depth 4 + prev_illum 8 + prev_spp 1; writes motion 4 + spp 1 + illum 8 + depth history 4
BYTES_BMFR = 37 + 12           # reads depth 4 + normal 8 + noisyAcc 8 + albedo 4 + motion 4 + spp 1 + history 8; writes denoised 8 + final 4
BYTES_TAA = 12 + 4             # reads denoised 4 + motion 4 + history 4; writes final 4
BYTES_CHAIN_FUSED = 78         # ideal fully fused chain, the figure BASELINE.md quotes
INPUT_BYTES = 4 + 8 + 4 + 16   # depth + normal + albedo + raw rgba32f per pixel (host -> device per frame)
DEFAULT_WORKLOAD = "bmfr_taa_4k"


def seq_index(f, R):
    """frame number -> resident frame: the sequence is walked forwards and backwards (0 .. R-1, R-2 .. 1, 0 ..), so that
    consecutive frames are always neighbours on the camera path -- cycling f % R would jump the camera back to the start
    every R frames (a full-frame disocclusion, and more reprojection displacement than a band's halo covers)"""
    if R < 2:
        return 0
    m = f % (2 * R - 2)
    return m if m < R else 2 * R - 2 - m


def config_of(name):
    """the workload description both arms print verbatim (the driver compares the two `config` objects)"""
    W, H, taa, desc = WORKLOADS[name]
    return {"workload": name, "description": desc, "width": W, "height": H, "block": 32, "taa": taa,
            "sequence": "deterministic synthetic G-buffer sequence (analytic scene, 1-spp noise, camera motion), frames at full size",
            "l2": "GPU arms: every frame's input planes are distinct resident buffers read once per step (inputs larger than L2)"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons during the timed region (NVML)"""

    NAMES = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz, self._halt = device, [], set(), None, threading.Event()
        self._nv = self._h = None
        try:        # NVML is initialised here, before the timed region: the region itself can be as short as 15 ms
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def sample(self):
        nv, h = self._nv, self._h
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, n in self.NAMES.items():
                if r & bit:
                    self.reasons.add(n)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def run(self):
        # 200 Hz: the shortest timed region (60 frames x 0.24 ms) still gets a few samples.  Not faster: NVML queries
        # go through the driver and visibly perturb kernel launches (measured: 1 kHz polling cost the polled rank
        # 0.09 ms/frame in the 2-GPU run, which its neighbours then wait for)
        while not self._halt.is_set():
            self.sample()
            self._halt.wait(0.005)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def render_sequence(W, H, n, rows=None, pinned=True, first=0):
    """n frames of the synthetic sequence in (pinned) host memory: dict of torch tensors + cameras"""
    import torch

    from vulkanpbrt_b200 import synth
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pinned)
    seq = {"depth": mk((n, H, W), torch.float32), "normal": mk((n, H, W, 2), torch.float32),
           "albedo": mk((n, H, W, 4), torch.uint8), "illum": mk((n, H, W, 4), torch.float32), "cams": []}
    material = np.zeros((H, W, 4), np.uint8)
    for i in range(n):
        fr = synth.Frame(first + i, seq["depth"][i].numpy(), seq["normal"][i].numpy(), seq["albedo"][i].numpy(), material,
                         seq["illum"][i].numpy(), None)
        synth.render_frame(W, H, first + i, rows=rows, out=fr)
        seq["cams"].append(fr.camera)
    return seq


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path
# ---------------------------------------------------------------------------------------------------
def cpu_chain(W, H, taa):
    """the reference's CPU implementation of the path: oracle/_ref (the reference's shader source compiled through
    oracle/glsl_shim) when it is built, else the oracle port"""
    from oracle import oracle as O
    from oracle import ref as R
    if R.available():
        return O, R.RefChain(W, H, "bmfr", 32, use_taa=taa), "reference"
    return O, O.OracleChain(W, H, "bmfr", 32, use_taa=taa), "port"


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms use every core of the box.  Must run before the
    oracle libraries are loaded (libgomp reads the variable once)."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def run_cpu(name, frames, budget_s=None, warmup=0):
    """times the CPU chain on `frames` synthetic frames of the workload at its REAL size; stops early once budget_s is
    spent (cpu_baseline leg: a bounded sample)"""
    use_all_host_cores()
    from vulkanpbrt_b200 import synth
    W, H, taa, _ = WORKLOADS[name]
    O, chain, kind = cpu_chain(W, H, taa)
    try:
        O.lib().vkpbrt_oracle_set_num_threads(os.cpu_count() or 1)
    except AttributeError:
        pass
    nf = max(1, min(frames + warmup, 4))
    fs = [synth.render_frame(W, H, f) for f in range(nf)]
    for f in range(warmup):
        chain.run_frame(f, fs[f % nf])
    t0 = time.perf_counter()
    done = 0
    for f in range(warmup, warmup + frames):
        chain.run_frame(f, fs[f % nf])
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    what = ("reference shader source (shaders/*.comp) compiled for the CPU via oracle/glsl_shim" if kind == "reference"
            else "oracle/vkpbrt_oracle.c")
    return {"value": W * H * done / dt / 1e6, "unit": "MPix/s", "cores": int(O.lib().vkpbrt_oracle_num_threads()),
            "kind": kind, "sample": f"{done} frames of {W}x{H} ({what}, OpenMP over workgroups), {dt:.1f} s",
            "ms_per_frame": dt / done * 1e3}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    r = run_cpu(name, args.steps, warmup=min(args.warmup, 2))
    line = {"impl": "reference", "metric": "BMFR denoised MPix/s", "value": r["value"], "unit": "MPix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(name),
            "note": "the reference's CPU path (no Vulkan ICD / lavapipe on this machine): " + r["sample"]
                    + f"; {min(args.warmup, 2)} untimed warm-up frames",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def measure_single(args, name, K, Wm, R, local, with_e2e=True, cpu_budget=0.0):
    """one workload on ONE GPU: resident-input throughput, per-kernel times, roofline, (optionally) the end-to-end leg"""
    import torch

    from vulkanpbrt_b200 import Context, DenoisePipeline, DenoisingBlockSize, DenoisingType

    W, H, taa, desc = WORKLOADS[name]
    dev = torch.device("cuda", local)
    # an explicit stream: the library records on the stream it is handed and the CUDA events below must sit on the same one
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = Context(local, stream.cuda_stream)
    assert ctx.stream == stream.cuda_stream
    t_gen = time.perf_counter()
    seq = render_sequence(W, H, R)
    t_gen = time.perf_counter() - t_gen
    dseq = {k: seq[k].to(dev, non_blocking=True) for k in ("depth", "normal", "albedo", "illum")}
    torch.cuda.synchronize()

    bfr = name.startswith("bfr")

    def make_pipe():
        if bfr:
            return DenoisePipeline(W, H, DenoisingType.BFR, DenoisingBlockSize.X8X16X32, use_taa=taa, ctx=ctx, external_inputs=True,
                                   average_squared=True)
        return DenoisePipeline(W, H, DenoisingType.BMFR, DenoisingBlockSize.X32, use_taa=taa, ctx=ctx, external_inputs=True)

    def bind(pipe, bufs, i):
        pipe.bind_inputs(bufs["depth"][i].data_ptr(), bufs["normal"][i].data_ptr(), bufs["albedo"][i].data_ptr(),
                         bufs["illum"][i].data_ptr())

    # ---- value: inputs resident in HBM -----------------------------------------------------------------
    pipe = make_pipe()

    def frame_resident(f):
        i = seq_index(f, R)
        bind(pipe, dseq, i)
        pipe.set_frame_constants(f, seq["cams"][i])
        pipe.record()
        pipe.end_frame(seq["cams"][i])

    for f in range(Wm):
        frame_resident(f)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_host = time.perf_counter()
    for f in range(Wm, Wm + K):
        frame_resident(f)
    e1.record(stream)
    t_host = (time.perf_counter() - t_host) / K * 1e3          # host time to ENQUEUE one frame (no sync inside)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    value = W * H * K / (ms * 1e-3) / 1e6

    # ---- per-kernel times (CUDA events around each recorded command), same steady state ---------------------
    ncmd = len(pipe.commands.children)
    names = pipe.command_labels
    probe_ms = [[] for _ in range(ncmd)]
    frame_lat = []
    nprobe = min(60, max(K, 40))
    lanes = [(m, l) for m, l in zip(pipe.modules, (0, 1, 2)) if bfr and hasattr(m, "set_lane")]
    for m, _ in lanes:
        m.set_lane(0)            # side lanes overlap the three block sizes: per-kernel event times only exist when they run in sequence
    for f in range(Wm + K, Wm + K + nprobe):
        i = seq_index(f, R)
        bind(pipe, dseq, i)
        pipe.set_frame_constants(f, seq["cams"][i])
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(ncmd + 1)]
        evs[0].record(stream)
        for c in range(ncmd):
            pipe.commands.children[c](pipe.commands)
            evs[c + 1].record(stream)
        pipe.end_frame(seq["cams"][i])
        torch.cuda.synchronize()
        for c in range(ncmd):
            probe_ms[c].append(evs[c].elapsed_time(evs[c + 1]))
        frame_lat.append(evs[0].elapsed_time(evs[ncmd]))
    for m, l in lanes:
        m.set_lane(l)
    acc_ms = [sorted(v)[len(v) // 2] for v in probe_ms]       # median over the probe frames: robust against a one-off hiccup
    frame_lat.sort()
    # one frame at a time (device idle before each): the latency a frame-by-frame caller sees (median, p95)
    latency = {"median": round(frame_lat[len(frame_lat) // 2], 5), "p95": round(frame_lat[min(len(frame_lat) - 1, int(0.95 * len(frame_lat)))], 5),
               "frames": nprobe}
    # sampled over the timed region AND the per-kernel probes (both run the same kernels back to back; the timed region
    # alone is only K x 0.7 ms long, two or three NVML polls)
    clocks = sampler.stop()
    hbm_peak, peak_src = peaks()
    per_kernel_bytes = {"k_accumulate": BYTES_ACCUMULATE, "k_bmfr_block<32,256>": BYTES_BMFR, "k_taa": BYTES_TAA,
                        "k_bfr_block<8>": BYTES_BMFR - 8, "k_bfr_block<16>": BYTES_BMFR - 8, "k_bfr_block<32>": BYTES_BMFR - 8,
                        "k_bfr_blend": 8 + 8 + 12 + 4}
    kernels = {}
    for n, t in zip(names, acc_ms):
        if n in per_kernel_bytes and t > 0:
            gbs = per_kernel_bytes[n] * W * H / (t * 1e-3) / 1e9
            kernels[n] = {"ms": round(t, 4), "share": round(t / sum(acc_ms), 3), "algorithmic_bytes_per_pixel": per_kernel_bytes[n],
                          "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 3)}
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(name, {}).get(dom)
    # FP32 work of the fit (4.05e5 FLOP per 32x32 block as FMAs; executed as separate mul/add)
    # blocks of the padded, jittered grid that hold at least one image pixel (the others are not fitted), mean over the 16 phases
    from vulkanpbrt_b200.multigpu import block_offset

    def fitted(frame):
        ox, oy = block_offset(32, frame)
        nx = sum(1 for bx in range(W // 32 + 2) if bx * 32 - ox < W and bx * 32 - ox + 32 > 0)
        ny = sum(1 for by in range(H // 32 + 2) if by * 32 - oy < H and by * 32 - oy + 32 > 0)
        return nx * ny
    nblocks = sum(fitted(f) for f in range(16)) / 16.0
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": round(kernels[dom]["achieved_gbs"] / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_kernel_bytes[dom] * W * H,
                "note": "k_bmfr_block is bound by instruction issue (batched Householder QR on the FP32 pipe, no tensor cores by "
                        "design), see `issue` and fit_gflops; the streaming kernels' HBM fractions are under 'kernels'",
                "fit_gflops": round(nblocks * 4.05e5 / (kernels[dom]["ms"] * 1e-3) / 1e9, 1) if "bmfr" in dom else None,
                "chain_fused_bytes_per_pixel": BYTES_CHAIN_FUSED,
                "chain_achieved_gbs": round(BYTES_CHAIN_FUSED * W * H / (ms / K * 1e-3) / 1e9, 1)}

    # the dominant kernel is bound by instruction issue, not by HBM: report that roofline too.  Warp instructions per
    # launch come from the committed ncu capture of this workload (like `traffic`); peak = SMs x 4 schedulers x clock.
    try:
        ip = ROOT / "profiles" / "roofline_instructions.json"
        inst = json.loads(ip.read_text()).get(name, {}).get(dom) if ip.exists() else None
        mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
        if inst and mhz:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak = sms * 4 * mhz * 1e6
            achieved = inst / (kernels[dom]["ms"] * 1e-3)
            roofline["issue"] = {"warp_instructions_per_launch": int(inst), "achieved_ginst_per_s": round(achieved / 1e9, 1),
                                 "peak_ginst_per_s": round(peak / 1e9, 1), "frac": round(achieved / peak, 3),
                                 "source": "smsp__inst_executed.sum of profiles/*_ncu_summary (same workload), SMs x 4 issue slots x SM clock"}
    except Exception as e:  # pragma: no cover -- never let a reporting extra break the line
        roofline["issue"] = {"error": type(e).__name__}

    res = {"workload": name, "value": round(value, 1), "unit": "MPix/s", "ms_per_step": round(ms / K, 5), "steps": K, "warmup": Wm,
           "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(t_host, 4), "frame_latency_ms": latency, "clocks": clocks,
           "roofline": roofline, "kernels": kernels, "resident_frames": R, "sequence_generation_s": round(t_gen, 1), "e2e": None,
           "cpu_baseline": None,
           "kernels_note": ("per-kernel times are taken with the three block sizes in sequence; the timed region overlaps them on "
                            "side lanes (ms_per_step < sum of the kernels)") if bfr else None}

    # ---- e2e: host buffers, H2D + D2H inside the timed region ------------------------------------------------
    del pipe
    if with_e2e:
        pipe = make_pipe()
        copy_stream = torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(dseq[k][0]) for k in ("depth", "normal", "albedo", "illum")} for _ in range(2)]
        out_host = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        final_view = lambda: torch.as_tensor(pipe.final, device=dev)

        def issue_copy(f):
            s = f % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                for k in ("depth", "normal", "albedo", "illum"):
                    dbuf[s][k].copy_(seq[k][seq_index(f, R)], non_blocking=True)
                copied[s].record(copy_stream)

        def frame_e2e(f):
            s = f % 2
            stream.wait_event(copied[s])
            pipe.bind_inputs(dbuf[s]["depth"].data_ptr(), dbuf[s]["normal"].data_ptr(), dbuf[s]["albedo"].data_ptr(),
                             dbuf[s]["illum"].data_ptr())
            pipe.set_frame_constants(f, seq["cams"][seq_index(f, R)])
            pipe.record()
            pipe.end_frame(seq["cams"][seq_index(f, R)])
            consumed[s].record(stream)
            out_host[s].copy_(final_view(), non_blocking=True)       # the step's result, read back every frame

        for s in range(2):
            consumed[s].record(stream)
        issue_copy(0)
        for f in range(Wm):
            issue_copy(f + 1)
            frame_e2e(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        for f in range(Wm, Wm + K):
            issue_copy(f + 1)
            frame_e2e(f)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        e2e_value = W * H * K / (e2e_ms * 1e-3) / 1e6
        res["e2e"] = {"value": round(e2e_value, 1), "unit": "MPix/s", "h2d_bytes_per_step": INPUT_BYTES * W * H,
                      "d2h_bytes_per_step": 4 * W * H, "ms_per_step": round(e2e_ms / K, 5), "wall_ms_per_step": round(wall_ms / K, 5),
                      "h2d_gb_per_s": round(INPUT_BYTES * W * H / (e2e_ms / K * 1e-3) / 1e9, 1)}
        del pipe, dbuf, out_host
    del dseq, seq
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if cpu_budget > 0 and not bfr:
        res["cpu_baseline"] = run_cpu(name, frames=64, budget_s=cpu_budget, warmup=1)
    return res


def also_entry(r):
    """the nested form of a secondary workload's result"""
    return {"config": config_of(r["workload"]), "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "steps": r["steps"],
            "warmup": r["warmup"], "kernels": r["kernels"], "kernels_note": r.get("kernels_note"), "roofline": {k: r["roofline"][k] for k in ("kernel", "achieved", "peak", "frac")},
            "gpu_launches": r["gpu_launches"], "e2e": r["e2e"], "clocks": r["clocks"]}


def main_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        # exactly ONE line on stdout (the JSON): library chatter such as NCCL's version banner goes to stderr
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from vulkanpbrt_b200.multigpu import bench_multi
        line = bench_multi(args, rank, world, local)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    name = args.workload or DEFAULT_WORKLOAD
    W, H, taa, desc = WORKLOADS[name]
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.resident_frames)        # frames kept resident; longer runs cycle through them
    if name == "bmfr_8k":
        R = min(R, 8)                            # 1 GB of input planes per frame
    r = measure_single(args, name, K, Wm, R, local, with_e2e=True, cpu_budget=args.cpu_budget)
    also = {}
    if not args.no_also and args.workload is None:
        # the other BASELINE configurations, shorter runs of the same measurement
        for other, k2, w2, r2 in (("bmfr_1080p", 20, 5, 25), ("bfr_blend_1080p", 10, 3, 13), ("bmfr_8k", 8, 3, 6)):
            try:
                also[other] = also_entry(measure_single(args, other, k2, w2, r2, local, with_e2e=(other != "bmfr_8k"), cpu_budget=0.0))
            except Exception as e:  # pragma: no cover -- a secondary workload must not cost the headline line
                also[other] = {"error": f"{type(e).__name__}: {e}"}
    bfr = name.startswith("bfr")
    cfg = config_of(name)
    line = {"metric": ("BFR+blend" if bfr else "BMFR") + " denoised MPix/s", "value": r["value"], "unit": "MPix/s", "n_gpus": 1, "steps": K, "warmup": Wm,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            # the --gpus series cuts this same frame into N bands (total work fixed), so the N = 1 point carries the series' label
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "run": {"resident_frames": r["resident_frames"], "sequence_generation_s": r["sequence_generation_s"],
                    "l2": f"inputs larger than L2: {r['resident_frames']} resident frames x {INPUT_BYTES * W * H / 1e6:.0f} MB, each read once per step"},
            "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "host_enqueue_ms_per_step": r["host_enqueue_ms_per_step"],
            "frame_latency_ms": r["frame_latency_ms"], "clocks": r["clocks"], "roofline": r["roofline"], "kernels": r["kernels"],
            "cpu_baseline": r["cpu_baseline"], "also": also}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--resident-frames", type=int, default=48)
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent sequences of the workload, one per GPU")
    ap.add_argument("--weak", action="store_true", help="N > 1: weak scaling (one band of the workload's height per GPU) instead of strong")
    ap.add_argument("--no-verify", dest="no_verify", action="store_true", help="N > 1: skip the banded == single-GPU comparison")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo rows through NVLink peer memory (k_halo_push) or NCCL send/recv groups")
    ap.add_argument("--host", default="native", choices=["native", "python"],
                    help="N > 1: the per-frame driver -- vkpbrt::BandedRank (C++, one C-ABI call per frame) or the Python BandedPipeline")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline (0 = skip)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads nested under 'also'")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    main()
