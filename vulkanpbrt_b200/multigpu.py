"""Band-sharded multi-GPU execution of the denoising chain.

The reference is single-device; this is new work.  One process per GPU (torch.distributed, NCCL over
NVLink).  The frame is partitioned by horizontal bands of whole BLOCK ROWS of the jittered BMFR grid:
rank g owns block rows [b_g, b_{g+1}).  Blocks are independent, so the only inter-GPU traffic is

  * the HISTORY HALO: rows of prev_depth / accumulated illumination / sample counts / denoised
    history / TAA history that the next frame's reprojection (bilinear, displacement <= D rows, REPEAT
    wrap at the image edge) reads outside the rank's own rows, sent by their canonical owner; and
  * one row of the tone-mapped denoiser output on each side for TAA's 3x3 neighbourhood.

Both are neighbour exchanges of a few rows (tens of KB .. a few MB): latency- not bandwidth-bound, so
they are batched into one NCCL group per exchange point.  Accumulate runs redundantly on the apron
rows a rank's blocks read through the jitter/mirror footprint instead of exchanging current-frame
planes.  Every rank keeps full-frame planes (the whole 8K working set is 3.5 GB) and only computes its
band; the producer's input planes are band-local buffers addressed through a virtual full-frame base.

BandPlan is pure integer geometry (tested on CPU); BandedPipeline drives one rank.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple

import numpy as np

from . import _capi as capi
from .modules import DenoisingBlockSize, DenoisingType
from .pipeline import DenoisePipeline

# bmfrGeneral.comp:36 -- the y component of ivec2(vec2(b, b) * pixelOffsets[f % 16])
_OFFSETS = [(.7, .85), (.95, .5), (.43, .76), (.97, .03), (.37, .58), (.03, .36), (.81, .46), (0, .78), (.36, -.08),
            (-.06, 0), (.95, .1), (.85, .61), (.06, .1), (.43, .16), (0, .5), (.73, .38)]


def block_offset(block: int, frame: int) -> Tuple[int, int]:
    ox, oy = _OFFSETS[frame % 16]
    return int(np.float32(block) * np.float32(ox)), int(np.float32(block) * np.float32(oy))


def _mirror(x: int, s: int) -> int:
    if x < 0:
        return -x - 1
    if x >= s:
        return 2 * s - x - 1
    return x


Rows = Tuple[int, int]   # [lo, hi)


def _clip(r: Rows, H: int) -> Rows:
    lo, hi = max(0, r[0]), min(H, r[1])
    return (lo, max(lo, hi))


def _intersect(a: Rows, b: Rows) -> Rows:
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, max(lo, hi))


@dataclass
class Transfer:
    src: int
    dst: int
    plane: str
    rows: Rows


class BandPlan:
    """Integer geometry of the band partition for a W x H frame on `world` ranks."""

    def __init__(self, width: int, height: int, world: int, block: int = 32, max_disp_rows: int = 24, taa: bool = False):
        self.W, self.H, self.N, self.b, self.D, self.taa = width, height, world, block, max_disp_rows, taa
        self.nby = height // block + 2
        self.brow = [round(g * self.nby / world) for g in range(world + 1)]
        # block rows at a band edge whose output a neighbour can ask for: displacement + jitter shift of the boundary
        self.edge_block_rows = -(-(max_disp_rows + 1 + block) // block)
        self.bottom_wrap_block_rows = self.nby - (height - 1) // block     # block rows that can hold image row H-1
        if any(self.brow[g + 1] - self.brow[g] < 2 * self.edge_block_rows for g in range(world)):
            raise ValueError("bands must be at least two edge regions high")

    def block_rows(self, g: int) -> Rows:
        return (self.brow[g], self.brow[g + 1])

    def boundary(self, g: int, frame: int) -> int:
        """first image row of rank g's canonical band at `frame` (rows written by its BMFR blocks)"""
        if g <= 0:
            return 0
        if g >= self.N:
            return self.H
        _, oy = block_offset(self.b, frame)
        return min(self.H, max(0, self.b * self.brow[g] - oy))

    def owned_rows(self, g: int, frame: int) -> Rows:
        return (self.boundary(g, frame), self.boundary(g + 1, frame))

    def accumulate_rows(self, g: int, frame: int) -> Rows:
        """image rows the rank's blocks read through the jitter + mirror footprint (+1 for TAA's stencil is
        not needed: TAA reads the denoiser output, not accumulate planes, outside its own rows)"""
        _, oy = block_offset(self.b, frame)
        lo_abs, hi_abs = self.b * self.brow[g] - oy, self.b * self.brow[g + 1] - oy - 1
        rows = [_mirror(lo_abs, self.H), _mirror(hi_abs, self.H), _mirror(max(lo_abs, 0), self.H), _mirror(min(hi_abs, self.H - 1), self.H)]
        lo, hi = min(rows), max(rows) + 1
        # TAA on the owned rows reads motion there: owned rows are inside [lo, hi) by construction
        o = self.owned_rows(g, frame)
        return _clip((min(lo, o[0]), max(hi, o[1])), self.H)

    def input_rows(self, g: int) -> Rows:
        """rows of the producer's planes the rank ever touches (all 16 jitter phases)"""
        lo = min(self.accumulate_rows(g, f)[0] for f in range(16))
        hi = max(self.accumulate_rows(g, f)[1] for f in range(16))
        return (lo, hi)

    # ---- halo requirements of frame `frame` (reads of the history written at frame - 1) -----------------
    def _with_disp(self, r: Rows) -> List[Rows]:
        lo, hi = max(0, r[0] - self.D - 1), min(self.H, r[1] + self.D + 1)
        out = [(lo, hi)]
        # REPEAT addressing: a tap at row -1 / H wraps to the opposite image edge
        if lo == 0 and hi < self.H:
            out.append((self.H - 1, self.H))
        if hi == self.H and lo > 0:
            out.append((0, 1))
        return out

    def history_transfers(self, frame_next: int) -> List[Transfer]:
        """rows every rank must receive before running `frame_next`, sent by the canonical owner of the row
        at frame_next - 1.  Planes: acc (prev_depth, prev_illu, prev_spp), denoised, taa."""
        f0 = frame_next - 1
        out: List[Transfer] = []
        for dst in range(self.N):
            have_acc = self.accumulate_rows(dst, f0)
            have_own = self.owned_rows(dst, f0)
            need = {"acc": self._with_disp(self.accumulate_rows(dst, frame_next)),
                    "denoised": self._with_disp(self.owned_rows(dst, frame_next))}
            if self.taa:
                need["taa"] = self._with_disp(self.owned_rows(dst, frame_next))
            for plane, ranges in need.items():
                have = have_acc if plane == "acc" else have_own
                for r in ranges:
                    for src in range(self.N):
                        if src == dst:
                            continue
                        part = _intersect(r, self.owned_rows(src, f0))
                        # drop what the receiver computed itself (identical values)
                        for piece in _subtract(part, have):
                            if piece[1] > piece[0]:
                                out.append(Transfer(src, dst, plane, piece))
        return out

    def stale_column_transfers(self, frame: int) -> List[Transfer]:
        """Frames with a negative x jitter (f = 9 mod 16: offset (-1, 0)) leave image column 0 UNWRITTEN: the
        reference keeps whatever the previous frames stored there.  The band boundary
        moves with the y jitter, so the rank that owns such a pixel may not be the one that wrote it last.
        After every frame each rank therefore mirrors column 0 of the rows it wrote inside the window the
        boundary can move in to its neighbour (planes: denoiser final, denoised layer just written)."""
        out: List[Transfer] = []
        for g in range(1, self.N):
            win = _clip((self.b * self.brow[g] - self.b, self.b * self.brow[g] + 3), self.H)
            for src, dst in ((g - 1, g), (g, g - 1)):
                rows = _intersect(win, self.owned_rows(src, frame))
                if rows[1] > rows[0]:
                    out.append(Transfer(src, dst, "final_col0", rows))
                    out.append(Transfer(src, dst, "denoised_col0", rows))
        return out

    def final_transfers(self, frame: int) -> List[Transfer]:
        """one row of the denoiser's tone-mapped output on each side of the owned rows, for TAA (taa.comp:66-83)"""
        out: List[Transfer] = []
        if not self.taa:
            return out
        for dst in range(self.N):
            lo, hi = self.owned_rows(dst, frame)
            for row in (lo - 1, hi):
                if 0 <= row < self.H:
                    for src in range(self.N):
                        o = self.owned_rows(src, frame)
                        if src != dst and o[0] <= row < o[1]:
                            out.append(Transfer(src, dst, "final", (row, row + 1)))
        return out


def _subtract(a: Rows, b: Rows) -> List[Rows]:
    """a \\ b for half-open row ranges"""
    if a[1] <= a[0]:
        return []
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    if hi <= lo:
        return [a]
    out = []
    if a[0] < lo:
        out.append((a[0], lo))
    if hi < a[1]:
        out.append((hi, a[1]))
    return out


class NcclDirect:
    """Thin ctypes binding of the NCCL library torch already loaded: one ncclGroup of sends / receives per
    exchange point, issued on a dedicated communication stream.  torch.distributed (backend nccl) stays the
    plumbing -- rendezvous, the unique-id broadcast, barriers -- but its Python P2P path costs ~100 us of
    host time per group, which is the whole budget of a 240 us frame."""

    def __init__(self, dist, rank: int, world: int, device):
        import ctypes as C

        import torch
        path = None
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl" in line:
                    path = line.split()[-1]
                    break
        if path is None:
            raise RuntimeError("libnccl is not loaded in this process (torch.distributed nccl backend required)")
        self.C, self.lib = C, C.CDLL(path)

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]
        uid = UniqueId()
        if rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self.comm = C.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        for fn in (self.lib.ncclSend, self.lib.ncclRecv):
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.stream = torch.cuda.Stream(device=device)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"NCCL error {rc}")

    def group(self, ops) -> None:
        """ops: [(is_send, device_ptr, nbytes, peer)] issued as one group on the communication stream"""
        lib, st = self.lib, self.stream.cuda_stream
        self._check(lib.ncclGroupStart())
        for is_send, ptr, n, peer in ops:
            self._check((lib.ncclSend if is_send else lib.ncclRecv)(ptr, n, 1, peer, self.comm, st))    # 1 = ncclUint8
        self._check(lib.ncclGroupEnd())

    def close(self) -> None:
        if self.comm:
            self.lib.ncclCommDestroy.argtypes = [self.C.c_void_p]
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


class PeerDirect:
    """Halo exchange over NVLink peer memory (include/vkpbrt_b200.h, "Band-sharded multi-GPU runs"): every rank
    maps the allocations it writes into (cudaIpc handles passed around through torch.distributed) and each exchange
    point is ONE k_halo_push launch on a communication stream: it stores the rows straight into the receivers' HBM
    and then publishes a sequence number to their flag words; the receiver's main stream runs k_halo_wait on its
    own flag words before the consuming kernel.  No host synchronisation, no staging copies, no rendezvous except
    for the end-of-frame exchange, whose receivers first tell the senders (a "ready" word) that their frame is over
    -- the point at which NCCL would have posted the receive.

    flag words of a rank: done[group][src] for group in A, B, F, then ready[dst]."""

    GROUPS = {"A": 0, "B": 1, "F": 2}

    def __init__(self, dist, rank: int, world: int, device, ctx, timeout_ms: int = 20000, comm_stream: int = None):
        """dist: anything with all_gather_object(out_list, obj) (torch.distributed on GPUs).  comm_stream: a cudaStream_t to
        use instead of creating a high-priority torch stream (the CPU tests pass 0)."""
        import ctypes as C

        from . import _capi as capi
        from .modules import DescriptorImage
        if world > capi.HALO_MAX_PEERS:
            raise ValueError(f"at most {capi.HALO_MAX_PEERS} ranks per node")
        self.C, self.capi, self.dist, self.rank, self.world, self.ctx = C, capi, dist, rank, world, ctx
        self.timeout_ms = timeout_ms
        self.device = device
        if comm_stream is None:
            import torch
            # high priority: an exchange kernel is launched while the main stream keeps all SMs busy (k_bmfr_block, or the
            # next frame's k_accumulate) and its few CTAs should be dispatched ahead of the main kernel's pending ones
            # (measured at N = 2: no difference either way -- the remaining ~25 us per banded frame are kernel boundaries)
            self.stream = torch.cuda.Stream(device=device, priority=-1)
            comm_stream = self.stream.cuda_stream
        else:
            self.stream = None
        self.flags = DescriptorImage.create(ctx, capi.FORMAT_R32_SFLOAT, max(16, 4 * world), 1)
        self.flags.compile()            # allocates, zero-initialised
        ctx.synchronize()
        self._exchanges: list = []
        lib = capi.lib()
        self._start_fn, self._wait_fn = lib.vkpbrt_halo_exchange_start, lib.vkpbrt_halo_exchange_wait
        self._comm = comm_stream
        self._opened: Dict = {}         # (rank, handle bytes) -> mapped base
        self.seq = {"A": 0, "B": 0, "F": 0}
        mine = self.export(self.flags.device_ptr)
        everyone = self.all_gather(mine)
        self.peer_flags = [self.flags.device_ptr if r == rank else self.map(r, everyone[r]) for r in range(world)]

    # ---- mappings ---------------------------------------------------------------------------------------
    def export(self, device_ptr: int):
        C = self.C
        h = (C.c_uint8 * self.capi.PEER_HANDLE_BYTES)()
        off = C.c_uint64()
        self.capi.call("vkpbrt_peer_export", self.ctx.handle, C.c_void_p(device_ptr), h, C.byref(off))
        return bytes(h), int(off.value)

    def map(self, rank: int, exported) -> int:
        C = self.C
        h, off = exported
        base = self._opened.get((rank, h))
        if base is None:
            out = C.c_void_p()
            buf = (C.c_uint8 * self.capi.PEER_HANDLE_BYTES).from_buffer_copy(h)
            self.capi.call("vkpbrt_peer_open", self.ctx.handle, buf, C.byref(out))
            base = self._opened[(rank, h)] = int(out.value)
        return base + off

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    # ---- flag addresses ---------------------------------------------------------------------------------
    def done_word(self, owner: int, group: str, src: int) -> int:
        return self.peer_flags[owner] + 4 * (self.GROUPS[group] * self.world + src)

    def ready_word(self, owner: int, dst: int) -> int:
        return self.peer_flags[owner] + 4 * (3 * self.world + dst)

    # ---- exchange points --------------------------------------------------------------------------------
    def make_exchange(self, kind: str, copies, announce, ready, done, wait):
        """copies: [(src, dst, src_pitch, dst_pitch, row_bytes, rows)]; the four flag lists are device addresses.
        Returns the opaque exchange handle (vkpbrt_halo_exchange_create); everything is resolved here, once."""
        C, capi = self.C, self.capi
        table = (capi.HaloCopy * max(1, len(copies)))()
        for k, c in enumerate(copies):
            table[k] = capi.HaloCopy(*c)
        arrays = [(C.c_void_p * max(1, len(lst)))(*lst) for lst in (announce, ready, done, wait)]
        d = capi.HaloExchangeDesc(table, len(copies), arrays[0], len(announce), arrays[1], len(ready), arrays[2], len(done),
                                  arrays[3], len(wait))
        h = C.c_void_p()
        capi.call("vkpbrt_halo_exchange_create", self.ctx.handle, C.byref(d), self.timeout_ms, C.byref(h))
        self._exchanges.append((kind, h))
        return h

    def start(self, x, after_stream: int, value: int) -> None:
        rc = self._start_fn(x, self._comm, after_stream, value)
        if rc:
            self.capi.check(rc)

    def wait(self, x, stream: int, value: int) -> None:
        rc = self._wait_fn(x, stream, value)
        if rc:
            self.capi.check(rc)

    def stats(self):
        """synchronises; {kind: (ms spent in the start kernels' gate, ms spent in the wait kernels)} and the error flag"""
        C = self.C
        out, err = {}, 0
        for kind, h in self._exchanges:
            g, w, e = C.c_uint64(), C.c_uint64(), C.c_uint32()
            self.capi.call("vkpbrt_halo_exchange_stats", h, C.byref(g), C.byref(w), C.byref(e))
            a = out.setdefault(kind, [0.0, 0.0])
            a[0] += g.value * 1e-6
            a[1] += w.value * 1e-6
            err |= e.value
        return out, err

    def check(self) -> None:
        """synchronises and raises if any flag wait timed out (a peer died or fell out of step)"""
        if self.stats()[1]:
            raise RuntimeError("halo exchange: a flag wait timed out (peer rank lost or out of step)")

    def close(self) -> None:
        for _, h in self._exchanges:
            self.capi.call("vkpbrt_halo_exchange_destroy", h)
        self._exchanges = []
        for base in self._opened.values():
            try:
                self.capi.call("vkpbrt_peer_close", self.ctx.handle, self.C.c_void_p(base))
            except Exception:
                pass
        self._opened = {}


class BandedPipeline:
    """One rank of the band-sharded chain.  `view(image)` must return a uint8 torch tensor aliasing the image's
    CURRENT device buffer as raw bytes, shaped [layers?][H][row bytes] (cuda_view() below on a GPU)."""

    def __init__(self, width: int, height: int, rank: int, world: int, use_taa: bool, ctx, view: Callable,
                 max_disp_rows: int = 24, external_inputs: bool = True, dist=None, nccl: "NcclDirect" = None,
                 peer: "PeerDirect" = None):
        self.rank, self.world, self.view = rank, world, view
        self.plan = BandPlan(width, height, world, 32, max_disp_rows, use_taa)
        self.pipe = DenoisePipeline(width, height, DenoisingType.BMFR, DenoisingBlockSize.X32, use_taa=use_taa, ctx=ctx,
                                    external_inputs=external_inputs)
        self.dist = dist
        self.nccl = nccl                      # direct NCCL issue path (GPU); None: torch.distributed P2P ops (gloo tests)
        self.peer = peer                      # NVLink peer-memory path (GPU, one node): takes precedence over nccl
        self.bmfr = self.pipe.modules[0]
        self.bmfr.set_block_row_range(*self.plan.block_rows(rank))
        if world > 1:
            # a rank holds history rows within D (+1) rows of the rows it computes: taps beyond that are counted, not
            # silently served from stale rows (check() raises)
            self.pipe.accumulator.set_max_displacement_rows(max_disp_rows)
        c = self.pipe.commands.children
        self._acc_cmd, self._bmfr_cmd = c[0], c[1]
        self._taa_cmd = c[2] if use_taa else None
        self._back_cmd = c[-1]
        self.bytes_exchanged = 0
        self._pending_a = self._pending_b = None
        self._main_stream = ctx.stream          # the stream the modules record on
        self._swaps = 0                         # copy_to_back pointer swaps so far: selects the ping-pong buffers
        self._views: Dict = {}
        self._desc: Dict = {}
        self._keep: list = []
        self._events: list = []
        self._ev_i = 0

    # ---- cached exchange descriptors -------------------------------------------------------------------
    # Row ranges repeat with the 16-frame jitter period and the ping-pong buffers with period 2, so the
    # P2POp lists (and the tensor views they alias) are built once per (phase, buffer pointers) and reused:
    # the per-frame host cost is one batch_isend_irecv per exchange point.
    def _view(self, img):
        i = img.info()
        key = (i.data, i.size_bytes)
        v = self._views.get(key)
        if v is None:
            v = self._views[key] = self.view(img)
        return v

    def _build(self, transfers: List[Transfer], planes: Dict[str, list]):
        dist = self.dist
        ops, scatter, gather, raw = [], [], [], []
        nbytes = 0
        for t in transfers:     # identical order on every rank
            if t.src != self.rank and t.dst != self.rank:
                continue
            for p, ncol_bytes in planes[t.plane]:
                sl = p[t.rows[0]:t.rows[1]] if ncol_bytes is None else p[t.rows[0]:t.rows[1], :ncol_bytes]
                if t.src == self.rank:
                    buf = sl
                    if ncol_bytes is not None:          # column strips are staged through a contiguous buffer
                        buf = sl.new_empty(sl.shape)
                        gather.append((buf, sl))
                    raw.append((True, buf.data_ptr(), buf.numel(), t.dst))
                    if self.nccl is None:
                        ops.append(dist.P2POp(dist.isend, buf, t.dst))
                else:
                    buf = sl
                    if ncol_bytes is not None:
                        buf = sl.new_empty(sl.shape)
                        scatter.append((sl, buf))
                    raw.append((False, buf.data_ptr(), buf.numel(), t.src))
                    if self.nccl is None:
                        ops.append(dist.P2POp(dist.irecv, buf, t.src))
                    nbytes += sl.numel() * sl.element_size()
                    self._keep.append(buf)
        return (ops, gather, scatter, nbytes, raw)

    def _build_peer(self, kind: str, transfers: List[Transfer], images: Dict[str, list]):
        """copy table of this rank's outgoing blocks of rows (device resident), the ranks it sends to / receives
        from.  Collective: every rank builds the same key at the same frame and contributes its image handles."""
        import torch
        pd, g = self.peer, self.rank
        entries = {}
        exported = []
        for name, lst in images.items():
            entries[name] = []
            for img, ncol in lst:
                im, layer = (img if isinstance(img, tuple) else (img, 0))
                i = im.info()
                entries[name].append((len(exported), i.data, layer * i.layer_pitch, i.row_pitch, ncol))
                exported.append(pd.export(i.data))
        everyone = pd.all_gather(exported)
        rows_of, send_to, recv_from, nbytes = [], set(), set(), 0
        for t in transfers:
            if t.src == t.dst:
                continue
            for idx, base, layer_off, pitch, ncol in entries[t.plane]:
                rb = pitch if ncol is None else ncol
                off = layer_off + t.rows[0] * pitch
                if t.src == g:
                    rows_of.append((base + off, pd.map(t.dst, everyone[t.dst][idx]) + off, pitch, pitch, rb, t.rows[1] - t.rows[0]))
                    send_to.add(t.dst)
                elif t.dst == g:
                    recv_from.add(t.src)
                    nbytes += rb * (t.rows[1] - t.rows[0])
        send_to, recv_from = sorted(send_to), sorted(recv_from)
        # the end-of-frame group rewrites texels the receiver still reads during its frame: receivers announce the end
        # of their frame to the senders, whose copy is gated on it (the point where NCCL would have posted the receive)
        handshake = kind == "B"
        x = pd.make_exchange(kind, rows_of,
                             announce=[pd.ready_word(s_, g) for s_ in recv_from] if handshake else [],
                             ready=[pd.ready_word(g, d_) for d_ in send_to] if handshake else [],
                             done=[pd.done_word(d_, kind, g) for d_ in send_to],
                             wait=[pd.done_word(g, kind, s_) for s_ in recv_from])
        return ("peer", kind, x, bool(send_to or recv_from), nbytes)

    def _exchange_desc(self, kind: str, frame: int, images_fn, transfers_fn):
        """images_fn() -> plane name -> [(DescriptorImage or (DescriptorImage, layer), ncol_bytes)]"""
        if self.peer is not None:
            # every buffer involved alternates with the copy_to_back swaps (and the denoised layer with the frame
            # parity): no per-frame queries; the addresses a cached entry was built for are re-checked on its first reuses
            key = (kind, frame % 16, self._swaps & 1, frame & 1)
            entry = self._desc.get(key)
            if entry is None or entry[1] < 2:
                images = images_fn()
                ptrs = tuple((img[0] if isinstance(img, tuple) else img).info().data for lst in images.values() for img, _ in lst)
                if entry is None:
                    entry = self._desc[key] = [self._build_peer(kind, transfers_fn(), images), 0, ptrs]
                else:
                    if entry[2] != ptrs:
                        raise RuntimeError("halo exchange: image buffers do not follow the ping-pong schedule the cache assumes")
                    entry[1] += 1
            return entry[0]
        images = images_fn()
        ptrs = tuple((img[0] if isinstance(img, tuple) else img).info().data for lst in images.values() for img, _ in lst)
        key = (kind, frame % 16, ptrs)
        d = self._desc.get(key)
        if d is None:
            planes = {}
            for name, lst in images.items():
                planes[name] = []
                for img, ncol in lst:
                    if isinstance(img, tuple):
                        planes[name].append((self._view(img[0])[img[1]], ncol))
                    else:
                        planes[name].append((self._view(img), ncol))
            d = self._desc[key] = self._build(transfers_fn(), planes)
        return d

    def _event(self):
        import torch
        self._ev_i = (self._ev_i + 1) % len(self._events) if self._events else 0
        if len(self._events) < 16:
            self._events.append(torch.cuda.Event())
            return self._events[-1]
        return self._events[self._ev_i]

    def _start(self, desc):
        """enqueues one NCCL group with every send / recv of the descriptor and returns the pending handle;
        nothing waits.  NCCL orders the group after the work already enqueued on the current stream, so a
        group started right after a kernel overlaps whatever is enqueued next."""
        if self.world == 1 or desc is None:
            return None
        if desc[0] == "peer":
            return self._start_peer(desc)
        ops, gather, scatter, nbytes, raw = desc
        if not raw:
            return None
        for buf, sl in gather:
            buf.copy_(sl)
        self.bytes_exchanged += nbytes
        if self.nccl is not None:
            import torch
            cur = torch.cuda.current_stream()
            ready = self._event()
            ready.record(cur)
            self.nccl.stream.wait_event(ready)
            self.nccl.group(raw)
            done = self._event()
            done.record(self.nccl.stream)
            return (done, scatter)
        return (self.dist.batch_isend_irecv(ops), scatter)

    def _start_peer(self, desc):
        """one k_halo_push on the communication stream, ordered after the work already on the main stream"""
        _, kind, x, active, nbytes = desc
        pd = self.peer
        pd.seq[kind] += 1
        if not active:
            return None
        value = pd.seq[kind]
        self.bytes_exchanged += nbytes
        pd.start(x, self._main_stream, value)
        return ("peer", x, value)

    def _finish(self, pending) -> None:
        """makes the current stream wait for a group started by _start (stream-side wait for NCCL)"""
        if pending is None:
            return
        if pending[0] == "peer":
            self.peer.wait(pending[1], self._main_stream, pending[2])
            return
        works, scatter = pending
        if isinstance(works, list):
            for w in works:
                w.wait()
        else:
            import torch
            torch.cuda.current_stream().wait_event(works)
        for sl, buf in scatter:
            sl.copy_(buf)

    def run_frame(self, frame: int, cam) -> None:
        """inputs must already be bound / uploaded for this frame.  Per frame and boundary:
             A  accumulate-plane halos (depth history, accumulated illumination, sample counts): started right
                after k_accumulate, overlaps k_bmfr_block, awaited before the next frame's k_accumulate;
             F  one row of the denoiser output for TAA's stencil (TAA configurations only, not overlapped);
             B  denoised / TAA history halos + the stale-column strip: started at the end of the frame,
                overlaps the next frame's k_accumulate, awaited before its k_bmfr_block."""
        p, plan, g = self.pipe, self.plan, self.rank
        acc = p.accumulation_buffer
        multi = self.world > 1
        p.set_frame_constants(frame, cam)
        p.accumulator.set_row_range(*plan.accumulate_rows(g, frame))
        self._finish(self._pending_a)
        self._acc_cmd(p.commands)
        if multi:
            # pre-swap handles: what k_accumulate just wrote becomes prev_depth / prev_illu / prev_spp at copy_to_back
            da = self._exchange_desc("A", frame,
                                     lambda: {"acc": [(acc.next_depth, None), (p.illumination_buffer.illumination_images[0], None),
                                                      (acc.spp, None)]},
                                     lambda: [t for t in plan.history_transfers(frame + 1) if t.plane == "acc"])
            self._pending_a = self._start(da)
        self._finish(self._pending_b)
        self._pending_b = None
        self._bmfr_cmd(p.commands)
        if p.taa is not None:
            o0, o1 = plan.owned_rows(g, frame)
            p.taa.set_row_range(o0, o1)
            if multi and o1 - o0 > 2:
                # the band's first / last row need one row of the neighbour's tone-mapped output (taa.comp:66-83): the
                # rows in between run while that row is in flight, the edge rows after it has landed
                df = self._exchange_desc("F", frame, lambda: {"final": [(p.denoiser_final, None)]}, lambda: plan.final_transfers(frame))
                pending_f = self._start(df)
                i0 = o0 + (1 if g > 0 else 0)
                i1 = o1 - (1 if g < self.world - 1 else 0)
                p.taa.record_part(p.push_constants, i0, i1, False)
                self._finish(pending_f)
                p.taa.record_parts(p.push_constants, o0, i0, i1, o1, True)     # both edge rows, one launch (either may be empty): hands final -> history
            else:
                if multi:
                    df = self._exchange_desc("F", frame, lambda: {"final": [(p.denoiser_final, None)]}, lambda: plan.final_transfers(frame))
                    self._finish(self._start(df))
                self._taa_cmd(p.commands)
        self._back_cmd(p.commands)
        p.end_frame(cam)
        self._swaps += 1
        if multi:
            layer = (frame & 1) ^ 1

            def images_b():
                im = {"denoised": [((self.bmfr.denoised, layer), None)],
                      "final_col0": [(p.denoiser_final, 4)],                     # 1 BGRA8 texel
                      "denoised_col0": [((self.bmfr.denoised, layer), 8)]}       # 1 rgba16f texel
                if p.taa is not None:
                    im["taa"] = [(p.taa.history, None)]
                return im
            db = self._exchange_desc("B", frame, images_b,
                                     lambda: [t for t in plan.history_transfers(frame + 1) if t.plane != "acc"]
                                     + plan.stale_column_transfers(frame))
            self._pending_b = self._start(db)

    def flush(self) -> None:
        """waits (stream-side) for the halos in flight; call before reading planes outside the owned rows"""
        self._finish(self._pending_a)
        self._finish(self._pending_b)
        self._pending_a = self._pending_b = None

    def check(self) -> None:
        """after a device synchronisation: raises if the peer-memory flag protocol reported a timeout, or if a
        reprojection left the history rows this rank holds (camera motion larger than the plan's max_disp_rows)"""
        if self.peer is not None:
            self.peer.check()
        if self.world > 1:
            n = self.pipe.accumulator.displacement_violations()
            if n:
                raise RuntimeError(f"band-sharded run: {n} reprojection taps moved more than max_disp_rows = {self.plan.D} rows; "
                                   "the halo does not cover this camera motion (rebuild the BandPlan with a larger max_disp_rows)")

    def owned_rows(self, frame: int) -> Rows:
        return self.plan.owned_rows(self.rank, frame)


class NativeBandedRank:
    """vkpbrt::BandedRank (include/vkpbrt/banded.hpp) through the C ABI: the native host of the band-sharded chain.  One
    C call per frame instead of ~15 ctypes calls and their Python glue -- at N = 8 the Python driver above spends more
    host time per frame (0.13 ms) than a rank's share of a 4K frame takes on its GPU (0.09 ms).
    dist: anything with all_gather_object (torch.distributed on GPUs, a thread group in the emulator tests)."""

    def __init__(self, width: int, height: int, rank: int, world: int, use_taa: bool, ctx, dist=None, max_disp_rows: int = 24,
                 external_inputs: bool = True, comm_stream: int = None, timeout_ms: int = 20000):
        import ctypes as C
        self.C, self.ctx, self.rank, self.world, self.dist = C, ctx, rank, world, dist
        self.width, self.height = width, height
        self._own_stream = None
        if comm_stream is None and world > 1:
            s = C.c_void_p()
            capi.call("vkpbrt_stream_create", ctx.handle, 1, C.byref(s))
            self._own_stream = comm_stream = s.value

        def all_gather(_user, mine, nbytes, everyone):
            try:
                out = [None] * world
                dist.all_gather_object(out, C.string_at(mine, nbytes))
                C.memmove(everyone, b"".join(out), nbytes * world)
                return 0
            except Exception:       # noqa: BLE001 -- reported as a C-ABI error by the library
                import traceback
                traceback.print_exc()
                return 1
        self._cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)(all_gather)
        self._h = C.c_void_p()
        capi.call("vkpbrt_banded_rank_create", ctx.handle, width, height, rank, world, 1 if use_taa else 0, int(max_disp_rows),
                  1 if external_inputs else 0, C.c_void_p(comm_stream or 0), C.cast(self._cb, C.c_void_p) if world > 1 else None, None,
                  int(timeout_ms), C.byref(self._h))
        lib = capi.lib()
        self._run, self._bind = lib.vkpbrt_banded_rank_run_frame, lib.vkpbrt_banded_rank_bind_inputs
        from .modules import DescriptorImage
        img = lambda which: (lambda h: (capi.call("vkpbrt_banded_rank_image", self._h, which, C.byref(h)), DescriptorImage(ctx, h, False))[1])(C.c_void_p())
        self.final, self.denoiser_final, self.denoised = img(0), img(1), img(2)
        b = (C.c_int * (world + 1))()
        capi.call("vkpbrt_banded_rank_block_rows", self._h, b)
        self.brow = list(b)
        self.max_disp_rows = max_disp_rows

    def input_rows(self) -> Rows:
        lo, hi = self.C.c_int(), self.C.c_int()
        capi.call("vkpbrt_banded_rank_input_rows", self._h, self.C.byref(lo), self.C.byref(hi))
        return (lo.value, hi.value)

    def owned_rows(self, frame: int) -> Rows:
        lo, hi = self.C.c_int(), self.C.c_int()
        capi.call("vkpbrt_banded_rank_owned_rows", self._h, frame, self.C.byref(lo), self.C.byref(hi))
        return (lo.value, hi.value)

    def bind_inputs(self, depth_ptr: int, normal_ptr: int, albedo_ptr: int, illumination_ptr: int) -> None:
        rc = self._bind(self._h, depth_ptr, normal_ptr, albedo_ptr, illumination_ptr)
        if rc:
            capi.check(rc)

    @staticmethod
    def camera_block(cam):
        """view, inv_view, proj, inv_proj as one contiguous float32[64] (build once per camera, reuse every frame)"""
        return np.ascontiguousarray(np.concatenate([np.asarray(m, np.float32).reshape(16) for m in (cam.view, cam.inv_view, cam.proj, cam.inv_proj)]))

    def run_frame(self, frame: int, cam_block: np.ndarray) -> None:
        rc = self._run(self._h, frame, cam_block.ctypes.data)
        if rc:
            capi.check(rc)

    def flush(self) -> None:
        capi.call("vkpbrt_banded_rank_flush", self._h)

    def check(self) -> None:
        capi.call("vkpbrt_banded_rank_check", self._h)

    def stats(self):
        """synchronises; ({group: [gate ms, wait ms]}, bytes pushed)"""
        s, b = (self.C.c_uint64 * 8)(), self.C.c_uint64()
        capi.call("vkpbrt_banded_rank_stats", self._h, s, self.C.byref(b))
        return {k: [s[2 * i] * 1e-6, s[2 * i + 1] * 1e-6] for i, k in enumerate(("A", "B", "C", "D"))}, int(b.value)

    def close(self) -> None:
        if self._h:
            capi.call("vkpbrt_banded_rank_destroy", self._h)
            self._h = None
        if self._own_stream:
            capi.call("vkpbrt_stream_destroy", self.ctx.handle, self.C.c_void_p(self._own_stream))
            self._own_stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cuda_view(device):
    """view(image) for CUDA ranks: zero-copy uint8 tensor over the image's current device buffer"""
    import torch

    def view(img):
        return torch.as_tensor(img.byte_view(), device=device)
    return view


# ---------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): strong scaling (the workload's frame cut into N bands); --weak: one 1080p band per GPU
# ---------------------------------------------------------------------------------------------------
def _sha(t) -> str:
    import hashlib
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def _progress(rank: int, what: str) -> None:
    """phase markers on stderr (rank 0): a crash in a multi-rank run leaves no Python traceback"""
    import sys
    if rank == 0:
        print(f"[bench_multi] {what}", file=sys.stderr, flush=True)


class _PythonAdapter:
    """BandedPipeline behind the small interface bench_multi drives"""

    def __init__(self, bp: "BandedPipeline"):
        self.bp = bp

    final = property(lambda self: self.bp.pipe.final)
    denoised = property(lambda self: self.bp.bmfr.denoised)

    def prepare_camera(self, cam):
        return cam

    def bind_inputs(self, *ptrs):
        self.bp.pipe.bind_inputs(*ptrs)

    def run_frame(self, f, cam):
        self.bp.run_frame(f, cam)

    def flush(self):
        self.bp.flush()

    def check(self):
        self.bp.check()

    def owned_rows(self, f):
        return self.bp.owned_rows(f)

    def spin_stats(self):
        return dict(self.bp.peer.stats()[0]) if self.bp.peer is not None else {}

    def bytes_exchanged(self):
        return self.bp.bytes_exchanged

    def close(self):
        if self.bp.peer is not None:
            self.bp.peer.close()


class _NativeAdapter:
    """NativeBandedRank behind the same interface"""

    def __init__(self, nr: "NativeBandedRank"):
        self.nr = nr
        self.final, self.denoised = nr.final, nr.denoised

    def prepare_camera(self, cam):
        return NativeBandedRank.camera_block(cam)

    def bind_inputs(self, *ptrs):
        self.nr.bind_inputs(*ptrs)

    def run_frame(self, f, cam):
        self.nr.run_frame(f, cam)

    def flush(self):
        self.nr.flush()

    def check(self):
        self.nr.check()

    def owned_rows(self, f):
        return self.nr.owned_rows(f)

    def spin_stats(self):
        return self.nr.stats()[0] if self.nr.world > 1 else {}

    def bytes_exchanged(self):
        return self.nr.stats()[1]

    def close(self):
        self.nr.close()


def measure_banded(args, name: str, K: int, Wm: int, R: int, rank: int, world: int, local: int, weak: bool, replicas: bool, with_e2e: bool):
    """one workload band-sharded over the ranks.  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    from . import synth
    from .modules import Context
    import bench as B     # the repo-root bench.py (constants, sampler)

    W, Hband, taa, desc = B.WORKLOADS[name]
    _progress(rank, f"{name}: start")
    strong = not weak and not replicas
    H = Hband if (strong or replicas) else Hband * world
    prank, pworld = (0, 1) if replicas else (rank, world)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = Context(local, stream.cuda_stream)
    assert ctx.stream == stream.cuda_stream
    view = cuda_view(dev)
    halo = getattr(args, "halo", "peer")
    host = getattr(args, "host", "native")
    # reprojection displacement grows with the resolution: 24 rows per 1080 rows of image height
    max_disp = max(24, -(-24 * Hband // 1080))
    plan = BandPlan(W, H, pworld, 32, max_disp, taa)
    native = host == "native" and halo == "peer"
    if native:
        # the C++ driver (include/vkpbrt/banded.hpp) through the C ABI: one call per frame
        nr = NativeBandedRank(W, H, prank, pworld, taa, ctx, dist=dist, max_disp_rows=max_disp, external_inputs=True)
        rank_obj = _NativeAdapter(nr)
    else:
        bp = BandedPipeline(W, H, prank, pworld, taa, ctx, view, max_disp_rows=max_disp, dist=dist,
                            nccl=NcclDirect(dist, rank, world, dev) if (halo == "nccl" and not replicas) else None,
                            peer=PeerDirect(dist, rank, world, dev, ctx) if (halo == "peer" and not replicas) else None)
        rank_obj = _PythonAdapter(bp)
    bp = rank_obj
    lo, hi = plan.input_rows(prank)
    rows = hi - lo
    # Parity evidence that travels with the number: rank 0 ALSO runs the same frames through a plain single-GPU pipeline
    # (outside every timed region) and the ranks' owned rows are compared with it after the set-up frames.
    verify = (not replicas) and not getattr(args, "no_verify", False)
    keep_full = verify and rank == 0
    # band-local resident sequence; the kernels index absolute rows through a virtual full-frame base pointer
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    host = {"depth": mk((R, rows, W), torch.float32), "normal": mk((R, rows, W, 2), torch.float32),
            "albedo": mk((R, rows, W, 4), torch.uint8), "illum": mk((R, rows, W, 4), torch.float32)}
    cams = []
    full = synth.Frame(0, np.zeros((H, W), np.float32), np.zeros((H, W, 2), np.float32), np.zeros((H, W, 4), np.uint8),
                       np.zeros((H, W, 4), np.uint8), np.zeros((H, W, 4), np.float32), None)
    dfull = {k: [] for k in host}
    t_gen = time.perf_counter()
    for i in range(R):
        synth.render_frame(W, H, i, rows=None if keep_full else (lo, hi), out=full)
        host["depth"][i].numpy()[...] = full.depth[lo:hi]
        host["normal"][i].numpy()[...] = full.normal[lo:hi]
        host["albedo"][i].numpy()[...] = full.albedo[lo:hi]
        host["illum"][i].numpy()[...] = full.illumination[lo:hi]
        if keep_full:
            for k, a in (("depth", full.depth), ("normal", full.normal), ("albedo", full.albedo), ("illum", full.illumination)):
                dfull[k].append(torch.from_numpy(a).to(dev))
        cams.append(full.camera)
    t_gen = time.perf_counter() - t_gen
    dseq = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()
    _progress(rank, "inputs resident")
    pitch = {"depth": 4 * W, "normal": 8 * W, "albedo": 4 * W, "illum": 16 * W}

    cam_arg = [bp.prepare_camera(c) for c in cams]

    def bind(bufs, i):
        bp.bind_inputs(*[bufs[k][i].data_ptr() - lo * pitch[k] for k in ("depth", "normal", "albedo", "illum")])

    def frame(f):
        i = B.seq_index(f, R)
        bind(dseq, i)
        bp.run_frame(f, cam_arg[i])

    # set-up, not warm-up: the exchange descriptors (row ranges x ping-pong buffers: period 32 frames) are built
    # -- and for the peer path the neighbours' allocations mapped -- the first time each one is needed
    PRE = 32
    for f in range(PRE):
        frame(f)
    bp.flush()
    torch.cuda.synchronize()
    bp.check()
    _progress(rank, "set-up frames done")
    # ---- banded == single (every jitter phase has been through the exchange twice by now) -----------------------
    equal = None
    if verify:
        layer = ((PRE - 1) & 1) ^ 1
        mine = bp.owned_rows(PRE - 1)
        fin = view(bp.final)                            # [H][W*4]
        den = view(bp.denoised)                         # [2][H][W*8]
        my_hash = (_sha(fin[mine[0]:mine[1]]), _sha(den[layer, mine[0]:mine[1]]))
        hashes = [None] * world
        dist.all_gather_object(hashes, my_hash)
        if rank == 0:
            from .pipeline import DenoisePipeline
            sctx = Context(local, stream.cuda_stream)
            sp = DenoisePipeline(W, H, DenoisingType.BMFR, DenoisingBlockSize.X32, use_taa=taa, ctx=sctx, external_inputs=True)
            for f in range(PRE):
                i = B.seq_index(f, R)
                sp.bind_inputs(*[dfull[k][i].data_ptr() for k in ("depth", "normal", "albedo", "illum")])
                sp.set_frame_constants(f, cams[i])
                sp.record()
                sp.end_frame(cams[i])
            torch.cuda.synchronize()
            sfin, sden = view(sp.final), view(sp.modules[0].denoised)
            equal = True
            for g in range(world):
                o = plan.owned_rows(g, PRE - 1)
                equal = equal and hashes[g] == (_sha(sfin[o[0]:o[1]]), _sha(sden[layer, o[0]:o[1]]))
            del sp, sfin, sden
        dfull.clear()
        torch.cuda.empty_cache()
    dist.barrier()
    _progress(rank, f"verified: {equal}")
    for f in range(PRE, PRE + Wm):
        frame(f)
    torch.cuda.synchronize()
    stats0 = bp.spin_stats()
    # rank 0's GPU is the one whose clocks are reported.  NVML is initialised BEFORE the barrier: a rank that enters
    # the timed region late makes its neighbours' clocks run while they wait for its halos
    sampler = B.ClockSampler(local) if (rank == 0 and not os.environ.get("VKPBRT_NO_CLOCK_SAMPLER")) else None
    dist.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = None
    if rank == 0 and os.environ.get("VKPBRT_PROFILE_HOST"):
        import cProfile
        prof = cProfile.Profile()
    e0.record(stream)
    t_host = time.perf_counter()
    if prof:
        prof.enable()
    for f in range(PRE + Wm, PRE + Wm + K):
        frame(f)
    if prof:
        prof.disable()
    bp.flush()
    e1.record(stream)
    t_host = (time.perf_counter() - t_host) / K * 1e3      # host time to ENQUEUE one frame (no sync inside)
    if prof:
        import io
        import pstats
        import sys
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(22)
        print(buf.getvalue(), file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    bp.check()
    dist.barrier()
    clocks = sampler.stop() if sampler else None
    _progress(rank, "timed region done")
    # time each rank's streams spent spinning on flag words inside the timed region, per exchange group:
    # [gate of the end-of-frame push, wait in front of the consumer]; rank 0 reports every rank's
    spin = {}
    for kind, (g_ms, w_ms) in bp.spin_stats().items():
        g0, w0 = stats0.get(kind, (0.0, 0.0))
        spin[kind] = [round((g_ms - g0) / K, 4), round((w_ms - w0) / K, 4)]
    spins = [None] * world
    dist.all_gather_object(spins, spin)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    per_rank_ms = [None] * world
    dist.all_gather_object(per_rank_ms, round(float(t.item()) / K, 5))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = ctx.launch_count - launches0
    jobs = world if replicas else 1
    value = jobs * W * H * K / (ms * 1e-3) / 1e6

    # ---- e2e: per-rank band uploaded from pinned host memory every frame, owned rows of the result read back ----
    e2e = None
    nframes = PRE + Wm + K
    if with_e2e:
        copy_stream = torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(dseq[k][0]) for k in host} for _ in range(2)]
        out_host = torch.empty((H, W * 4), dtype=torch.uint8, pin_memory=True)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def issue_copy(f):
            s = f % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                for k in host:
                    dbuf[s][k].copy_(host[k][B.seq_index(f, R)], non_blocking=True)
                copied[s].record(copy_stream)

        def frame_e2e(f):
            s = f % 2
            stream.wait_event(copied[s])
            bp.bind_inputs(*[dbuf[s][k].data_ptr() - lo * pitch[k] for k in ("depth", "normal", "albedo", "illum")])
            bp.run_frame(f, cam_arg[B.seq_index(f, R)])
            consumed[s].record(stream)
            o = bp.owned_rows(f)
            out_host[o[0]:o[1]].copy_(view(bp.final)[o[0]:o[1]], non_blocking=True)     # [H][W*4] bytes

        for s in range(2):
            consumed[s].record(stream)
        f0 = PRE + Wm + K
        issue_copy(f0)
        for f in range(f0, f0 + 3):
            issue_copy(f + 1)
            frame_e2e(f)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record(stream)
        for f in range(f0 + 3, f0 + 3 + K):
            issue_copy(f + 1)
            frame_e2e(f)
        bp.flush()
        e1.record(stream)
        torch.cuda.synchronize()
        bp.check()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        nframes += 3 + K
        e2e = {"value": round(jobs * W * H * K / (e2e_ms * 1e-3) / 1e6, 1), "unit": "MPix/s",
               "h2d_bytes_per_step": B.INPUT_BYTES * W * rows * world, "d2h_bytes_per_step": 4 * W * H * jobs,
               "ms_per_step": round(e2e_ms / K, 5)}
        del dbuf, out_host
    _progress(rank, "e2e done")
    halo_bytes = torch.tensor([bp.bytes_exchanged()], device=dev, dtype=torch.float64)
    dist.all_reduce(halo_bytes)
    res = None
    if rank == 0:
        hbm_peak, peak_src = B.peaks()
        chain = (B.BYTES_ACCUMULATE + B.BYTES_BMFR + (B.BYTES_TAA if taa else 0))
        gbs = jobs * chain * W * H / (ms / K * 1e-3) / 1e9
        res = {"workload": name, "value": round(value, 1), "unit": "MPix/s", "ms_per_step": round(ms / K, 5), "steps": K, "warmup": Wm,
               "scaling": "strong" if strong else "weak", "replicas": jobs if replicas else None,
               "width": W, "height": H,
               "sharding": {"mode": ("replicas" if replicas else ("strong" if strong else "weak")) ,
                            "description": (f"{world} independent sequences, one per GPU, no communication" if replicas
                                            else f"frame {W}x{H} cut into {world} horizontal block bands"
                                            + ("" if strong else f" (weak scaling: one {W}x{Hband} band per GPU)")),
                            "band_block_rows": plan.brow,
                            "host": ("C++ (vkpbrt::BandedRank through the C ABI, one call per frame)" if native
                                     else "Python (BandedPipeline: one ctypes call per kernel / exchange point)"),
                            "halo": f"history rows +-{plan.D + 1} (+ REPEAT wrap row) per boundary, "
                                    + ("stored into the neighbours' HBM over NVLink peer mappings by k_halo_push, flag words for ordering"
                                       if halo == "peer" else "NCCL send/recv groups per frame"),
                            "resident_band_frames": R, "sequence_generation_s": round(t_gen, 1)},
               "banded_equals_single": equal,
               "banded_equals_single_note": (f"after {PRE} frames (every jitter phase twice): each rank's owned rows of the final image and of the "
                                             "denoised history, SHA-256, against a single-GPU run of the same frames on rank 0" if verify else None),
               "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
               "roofline": {"kernel": "chain (k_accumulate + k_bmfr_block" + (" + k_taa)" if taa else ")"), "bound": "hbm",
                            "achieved": round(gbs, 1), "peak": hbm_peak * world, "unit": "GB/s",
                            "frac": round(gbs / (hbm_peak * world), 4), "traffic": None, "peak_source": peak_src,
                            "note": "aggregate over ranks; per-kernel fractions are reported by the N=1 run"},
               "halo_bytes_per_step": float(halo_bytes.item()) / nframes, "host_enqueue_ms_per_step": round(t_host, 4),
               "halo_spin_ms_per_step_by_rank": spins, "ms_per_step_by_rank": per_rank_ms}
    dist.barrier()
    bp.close()
    del bp, rank_obj, dseq, host
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def bench_multi(args, rank: int, world: int, local: int):
    import torch.distributed as dist

    import bench as B

    name = args.workload or B.DEFAULT_WORKLOAD
    replicas = bool(getattr(args, "replicas", False))     # N independent sequences, one per GPU: no communication
    weak = bool(getattr(args, "weak", False))
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.resident_frames)
    if name == "bmfr_8k":
        R = min(R, 8)
    r = measure_banded(args, name, K, Wm, R, rank, world, local, weak, replicas, with_e2e=True)
    also = {}
    if world == 8 and args.workload is None and not getattr(args, "no_also", False) and not replicas and not weak:
        # BASELINE configs[4]: the 8K frame across the 8 GPUs
        try:
            r8 = measure_banded(args, "bmfr_8k", 10, 3, 6, rank, world, local, False, False, with_e2e=False)
            if r8 is not None:
                also["bmfr_8k"] = {"config": B.config_of("bmfr_8k"), **{k: r8[k] for k in
                                   ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "sharding", "banded_equals_single",
                                    "roofline", "halo_spin_ms_per_step_by_rank", "ms_per_step_by_rank", "gpu_launches")}}
        except Exception as e:  # pragma: no cover -- a secondary workload must not cost the headline line
            also["bmfr_8k"] = {"error": f"{type(e).__name__}: {e}"}
    line = None
    if rank == 0:
        cfg = B.config_of(name)
        if r["height"] != cfg["height"]:
            cfg = dict(cfg, workload=name + "_bands", height=r["height"])
        line = {"metric": "BMFR denoised MPix/s", "value": r["value"], "unit": "MPix/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"], "replicas": r["replicas"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "sharding": r["sharding"],
                "banded_equals_single": r["banded_equals_single"], "banded_equals_single_note": r["banded_equals_single_note"],
                "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "roofline": r["roofline"],
                "halo_bytes_per_step": r["halo_bytes_per_step"], "host_enqueue_ms_per_step": r["host_enqueue_ms_per_step"],
                "halo_spin_ms_per_step_by_rank": r["halo_spin_ms_per_step_by_rank"], "ms_per_step_by_rank": r["ms_per_step_by_rank"],
                "cpu_baseline": None, "also": also}
    dist.barrier()
    dist.destroy_process_group()
    return line
