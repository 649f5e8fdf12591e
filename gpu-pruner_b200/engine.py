"""Thin object wrapper over the C ABI (include/gpr.h) for Python callers, tests and bench.py.

The product is libgpr.so; this file only marshals arguments.  It mirrors the seam of the
reference at ``/root/reference/gpu-pruner/src/main.rs:397-437`` ("run the aggregation, get
back the candidate set"): :meth:`IdleEngine.decide` takes the window matrix and returns the
packed decision bitmap plus the ``QueryResponse``-style counts (``lib.rs:131-134``).

There is no CPU fallback anywhere in this module: if libgpr.so or a CUDA device is missing the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import ffi

_KERNELS = {"auto": ffi.GPR_KERNEL_AUTO, "ldg": ffi.GPR_KERNEL_LDG, "tma": ffi.GPR_KERNEL_TMA}


class GprError(RuntimeError):
    """Non-zero status from libgpr.so (the Rust wrapper would turn this into ``anyhow!``,
    feeding the failure counter at main.rs:310-321)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ffi.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


@dataclass
class Decision:
    n_pods: int
    decision_bits: np.ndarray            # uint32[ceil(P/32)] (x world with a communicator)
    candidate_bits: Optional[np.ndarray]
    series_max: Optional[np.ndarray]     # float32[P, G]
    n_series: int                        # QueryResponse.num_pods (series, pre-dedup; main.rs:418)
    n_candidates: int
    n_decisions: int
    kernel_ms: float
    veto_bits: Optional[np.ndarray] = None   # uint32[ceil(P/32)]: pods vetoed by the power clause (this rank's pods)

    def pods(self, bits: Optional[np.ndarray] = None) -> np.ndarray:
        """Indices of set bits (the idle-pod set), ascending."""
        w = self.decision_bits if bits is None else bits
        flat = np.unpackbits(np.ascontiguousarray(w, dtype="<u4").view(np.uint8), bitorder="little")
        return np.flatnonzero(flat)


def to_biased_u8(util: np.ndarray) -> np.ndarray:
    """f32 window (NaN = no sample) -> GPR_FMT_U8B bytes (0 = no sample, b = value + 1).  Raises if a
    sample is not an integer in 0..254 (DCGM_FI_DEV_GPU_UTIL is an integer percentage)."""
    util = np.asarray(util, dtype=np.float32)
    present = ~np.isnan(util)
    v = np.where(present, util, 0.0)
    if not np.all((v >= 0) & (v <= 254) & (v == np.floor(v))):
        raise ValueError("window is not representable in GPR_FMT_U8B (integers 0..254 or NaN)")
    return np.where(present, v + 1, 0).astype(np.uint8)


def from_biased_u8(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8)
    return np.where(b == 0, np.float32("nan"), b.astype(np.float32) - 1).astype(np.float32)


def _ptr(x) -> Optional[int]:
    """numpy array / torch tensor / int address / None -> address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(f"cannot take the address of {type(x)!r}")


class IdleEngine:
    """One context = one GPU.  Not re-entrant (one call at a time), thread-agnostic."""

    def __init__(self, device: int = 0, max_pods: int = 0, max_gpus: int = 0, max_samples: int = 0,
                 power_plane: bool = False, kernel: str = "auto", stream: Optional[int] = None):
        self._lib = ffi.load()
        cfg = ffi.gpr_config()
        cfg.struct_size = C.sizeof(ffi.gpr_config)
        cfg.device = device
        cfg.max_pods, cfg.max_gpus, cfg.max_samples = max_pods, max_gpus, max_samples
        cfg.flags = ffi.GPR_F_POWER_PLANE if power_plane else 0
        cfg.kernel_variant = _KERNELS[kernel]
        cfg.stream = stream
        h = C.c_void_p()
        rc = self._lib.gpr_create(C.byref(cfg), C.byref(h))
        if rc != ffi.GPR_OK:
            raise GprError(rc, (self._lib.gpr_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._keep = []  # result structs / arrays referenced by outstanding async calls
        self._host_arrays = []  # pinned allocations handed out by host_array(), freed in close()

    # ---- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != ffi.GPR_OK:
            raise GprError(rc, (self._lib.gpr_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            for ptr in getattr(self, "_host_arrays", []):
                self._lib.gpr_host_free(self._h, ptr)
            self._host_arrays = []
            self._lib.gpr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    # ---- hot path ---------------------------------------------------------------------------
    def _window(self, util, power, eligible, created_ts, cutoff_ts, P, G, T, row_stride,
                power_threshold, mem_kind, util_format: int = ffi.GPR_FMT_F32) -> ffi.gpr_window:
        w = ffi.gpr_window()
        w.util_format = util_format
        w.struct_size = C.sizeof(ffi.gpr_window)
        w.mem_kind = mem_kind
        w.util, w.power = _ptr(util), _ptr(power)
        w.eligible, w.created_ts = _ptr(eligible), _ptr(created_ts)
        w.cutoff_ts = int(cutoff_ts)
        w.n_pods, w.n_gpus, w.n_samples = P, G, T
        w.row_stride = row_stride
        w.power_threshold = float(power_threshold) if power_threshold is not None else 0.0
        return w

    def decide(self, util: np.ndarray, power: Optional[np.ndarray] = None,
               eligible: Optional[np.ndarray] = None, created_ts: Optional[np.ndarray] = None,
               cutoff_ts: int = 0, power_threshold: Optional[float] = 0.0,
               want_candidates: bool = True, want_series_max: bool = False,
               world: int = 1, want_veto: bool = False) -> Decision:
        """Blocking decision over a HOST window ``util[P, G, T]`` (float32, NaN = no sample; or
        uint8 in the biased byte format GPR_FMT_U8B, see :func:`to_biased_u8`)."""
        fmt = ffi.GPR_FMT_U8B if getattr(util, "dtype", None) == np.uint8 else ffi.GPR_FMT_F32
        util = np.ascontiguousarray(util, dtype=np.uint8 if fmt else np.float32)
        if util.ndim != 3:
            raise ValueError("util must be [pods, gpus, samples]")
        P, G, T = util.shape
        if power is not None:
            power = np.ascontiguousarray(power, dtype=np.float32)
            if power.shape != util.shape:
                raise ValueError("power must have util's shape")
        if eligible is not None:
            eligible = np.ascontiguousarray(eligible, dtype=np.uint8)
        if created_ts is not None:
            created_ts = np.ascontiguousarray(created_ts, dtype=np.int64)
        w = self._window(util, power, eligible, created_ts, cutoff_ts, P, G, T, 0,
                         power_threshold, ffi.GPR_MEM_HOST, fmt)
        W = (P + 31) // 32 * world
        dbits = np.zeros(max(W, 1), dtype=np.uint32)
        cbits = np.zeros(max(W, 1), dtype=np.uint32) if want_candidates else None
        smax = np.zeros((P, G), dtype=np.float32) if want_series_max else None
        vbits = np.zeros(max((P + 31) // 32, 1), dtype=np.uint32) if want_veto else None
        r = ffi.gpr_result()
        r.struct_size = C.sizeof(ffi.gpr_result)
        r.out_mem_kind = ffi.GPR_MEM_HOST
        r.decision_bits, r.candidate_bits, r.series_max = _ptr(dbits), _ptr(cbits), _ptr(smax)
        r.veto_bits = _ptr(vbits)
        self._check(self._lib.gpr_decide(self._h, C.byref(w), C.byref(r)))
        d = Decision(P, dbits[:W], None if cbits is None else cbits[:W], smax, r.n_series,
                     r.n_candidates, r.n_decisions, r.kernel_ms)
        d.veto_bits = None if vbits is None else vbits[:(P + 31) // 32]
        return d

    def decide_ptr(self, util, P: int, G: int, T: int, decision_bits, *, power=None, eligible=None,
                   created_ts=None, cutoff_ts: int = 0, power_threshold: Optional[float] = 0.0,
                   candidate_bits=None, series_max=None, veto_bits=None, row_stride: int = 0,
                   in_kind: int = ffi.GPR_MEM_DEVICE, out_kind: int = ffi.GPR_MEM_DEVICE,
                   blocking: bool = True, resident: bool = False,
                   util_format: int = ffi.GPR_FMT_F32) -> ffi.gpr_result:
        """Raw-pointer form (device tensors, pinned host buffers).  With ``blocking=False`` the
        call only enqueues; counters in the returned struct are valid after :meth:`sync`."""
        w = self._window(util, power, eligible, created_ts, cutoff_ts, P, G, T, row_stride,
                         power_threshold, in_kind, util_format)
        r = ffi.gpr_result()
        r.struct_size = C.sizeof(ffi.gpr_result)
        r.out_mem_kind = out_kind
        r.decision_bits, r.candidate_bits, r.series_max = (_ptr(decision_bits), _ptr(candidate_bits),
                                                           _ptr(series_max))
        r.veto_bits = _ptr(veto_bits)
        if resident:
            self._check(self._lib.gpr_decide_resident(self._h, C.byref(w), C.byref(r)))
        elif blocking:
            self._check(self._lib.gpr_decide(self._h, C.byref(w), C.byref(r)))
        else:
            self._keep.append((w, r))
            self._check(self._lib.gpr_decide_async(self._h, C.byref(w), C.byref(r)))
        return r

    def make_batch(self, calls):
        """Pre-marshal a list of decide_ptr-style keyword dicts into contiguous gpr_window /
        gpr_result arrays for :meth:`decide_batch_async` (build once, enqueue many times)."""
        n = len(calls)
        wins = (ffi.gpr_window * n)()
        ress = (ffi.gpr_result * n)()
        for i, kw in enumerate(calls):
            w = self._window(kw["util"], kw.get("power"), kw.get("eligible"), kw.get("created_ts"),
                             kw.get("cutoff_ts", 0), kw["P"], kw["G"], kw["T"], kw.get("row_stride", 0),
                             kw.get("power_threshold", 0.0), kw.get("in_kind", ffi.GPR_MEM_DEVICE),
                             kw.get("util_format", ffi.GPR_FMT_F32))
            C.memmove(C.byref(wins, i * C.sizeof(ffi.gpr_window)), C.byref(w), C.sizeof(ffi.gpr_window))
            r = ress[i]
            r.struct_size = C.sizeof(ffi.gpr_result)
            r.out_mem_kind = kw.get("out_kind", ffi.GPR_MEM_DEVICE)
            r.decision_bits = _ptr(kw["decision_bits"])
            r.candidate_bits = _ptr(kw.get("candidate_bits"))
            r.series_max = _ptr(kw.get("series_max"))
        return wins, ress, calls   # `calls` keeps the tensors alive

    def decide_batch_async(self, batch, n: Optional[int] = None):
        wins, ress, _ = batch
        n = len(wins) if n is None else n
        self._check(self._lib.gpr_decide_batch_async(self._h, wins, ress, n))
        return ress

    def sync(self):
        self._check(self._lib.gpr_sync(self._h))
        self._keep.clear()

    # ---- resident window (daemon mode) --------------------------------------------------------
    def resident_init(self, P: int, G: int, T: int, power_plane: bool = False, block_index: bool = False):
        flags = (ffi.GPR_F_POWER_PLANE if power_plane else 0) | (ffi.GPR_F_BLOCK_INDEX if block_index else 0)
        self._check(self._lib.gpr_resident_init(self._h, P, G, T, flags))

    def resident_reindex(self):
        self._check(self._lib.gpr_resident_reindex(self._h))

    def append(self, util_cols, power_cols=None, n_new: Optional[int] = None, row_stride: int = 0,
               mem_kind: int = ffi.GPR_MEM_HOST):
        if isinstance(util_cols, np.ndarray):
            util_cols = np.ascontiguousarray(util_cols, dtype=np.float32)
            n_new = util_cols.shape[-1]
            if power_cols is not None:
                power_cols = np.ascontiguousarray(power_cols, dtype=np.float32)
        self._check(self._lib.gpr_append(self._h, _ptr(util_cols), _ptr(power_cols), n_new,
                                         row_stride, mem_kind))  # blocking: the columns are consumed

    def resident_planes(self):
        u, p, ld = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self._lib.gpr_resident_planes(self._h, C.byref(u), C.byref(p), C.byref(ld)))
        return u.value, p.value, ld.value

    # ---- multi-GPU ------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(ffi.GPR_UNIQUE_ID_BYTES)
        rc = self._lib.gpr_comm_unique_id(buf)
        if rc != ffi.GPR_OK:
            raise GprError(rc, (self._lib.gpr_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, ffi.GPR_UNIQUE_ID_BYTES)
        self._check(self._lib.gpr_comm_init(self._h, buf, rank, world))

    def p2p_init(self, rank: int, world: int, max_pods_per_rank: int) -> bytes:
        """allocate this rank's exchange block; returns the 64-byte CUDA IPC handle to publish"""
        buf = C.create_string_buffer(ffi.GPR_P2P_HANDLE_BYTES)
        self._check(self._lib.gpr_p2p_init(self._h, rank, world, max_pods_per_rank, buf))
        return buf.raw

    def p2p_attach(self, handles):
        """handles: the world handles in rank order (bytes each); enables the fused exchange"""
        blob = b"".join(handles)
        self._check(self._lib.gpr_p2p_attach(self._h, C.create_string_buffer(blob, len(blob))))

    def comm_destroy(self):
        self._check(self._lib.gpr_comm_destroy(self._h))

    # ---- memory / measurement helpers -------------------------------------------------------------
    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.gpr_host_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int):
        self._check(self._lib.gpr_host_free(self._h, ptr))

    def host_array(self, shape, dtype) -> np.ndarray:
        """numpy view over freshly allocated PINNED host memory (freed with the engine)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self.host_alloc(max(n, 1))
        self._host_arrays.append(ptr)
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.gpr_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        self._check(self._lib.gpr_device_free(self._h, ptr))

    def memcpy(self, dst, src, nbytes: int, dst_kind: int, src_kind: int):
        self._check(self._lib.gpr_memcpy(self._h, _ptr(dst), _ptr(src), nbytes, dst_kind, src_kind))

    def timer_begin(self):
        self._check(self._lib.gpr_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_double()
        self._check(self._lib.gpr_timer_end(self._h, C.byref(ms)))
        return ms.value

    def step_stamps(self):
        """(begin_ns, ns[]) — device %globaltimer at the last timer_begin and at the completion of every
        decision retired by the last sync / blocking call"""
        n, t0 = C.c_uint32(), C.c_uint64()
        self._check(self._lib.gpr_step_stamps(self._h, None, 0, C.byref(n), C.byref(t0)))
        out = np.zeros(max(n.value, 1), np.uint64)
        self._check(self._lib.gpr_step_stamps(self._h, _ptr(out), n.value, C.byref(n), C.byref(t0)))
        return int(t0.value), out[:n.value]

    def phase_stamps(self) -> np.ndarray:
        """[n, 4] ns: fold start, folded, flags raised, peers arrived — per decision retired by the last sync"""
        n = C.c_uint32()
        self._check(self._lib.gpr_phase_stamps(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 4), np.uint64)
        self._check(self._lib.gpr_phase_stamps(self._h, _ptr(out), n.value, C.byref(n)))
        return out[:n.value].reshape(-1, 4)

    def p2p_debug(self, mode: int):
        self._check(self._lib.gpr_p2p_debug(self._h, mode))

    def flush_l2(self):
        self._check(self._lib.gpr_flush_l2(self._h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.gpr_launch_count(self._h, C.byref(n)))
        return n.value

    def device_info(self) -> dict:
        info = ffi.gpr_device_info()
        info.struct_size = C.sizeof(ffi.gpr_device_info)
        self._check(self._lib.gpr_get_device_info(self._h, C.byref(info)))
        return {"name": info.name.decode(errors="replace"), "sm_count": info.sm_count,
                "cc": (info.cc_major, info.cc_minor), "l2_bytes": info.l2_bytes,
                "hbm_bytes": info.hbm_bytes}

    # ---- device-side ingest of the response text ------------------------------------------------
    SPAN_DTYPE = np.dtype([("begin", "<u8"), ("end", "<u8"), ("row", "<u4"), ("flags", "<u4"),
                           ("n_in", "<u4"), ("n_oow", "<u4"), ("n_tiny", "<u4"), ("reserved", "<u4")])

    def text_scan(self, text, slot: int = 0, n_bytes: Optional[int] = None, mem_kind: int = ffi.GPR_MEM_HOST):
        """Upload response text (bytes, or a pinned uint8 array from :meth:`host_array`) into ``slot`` and
        return the sorted offsets of every ``},"values":[`` and ``"]]``."""
        if isinstance(text, (bytes, bytearray)):
            buf = np.frombuffer(text, dtype=np.uint8)
        else:
            buf = text
        n = int(buf.size if n_bytes is None else n_bytes)
        cap = max(1024, n // 256)
        while True:
            opens = np.empty(cap, np.uint64)
            closes = np.empty(cap, np.uint64)
            no, nc = C.c_uint64(0), C.c_uint64(0)
            rc = self._lib.gpr_text_scan(self._h, slot, _ptr(buf), n, mem_kind, _ptr(opens), _ptr(closes), cap,
                                         C.byref(no), C.byref(nc))
            if rc == ffi.GPR_E_CAPACITY:
                cap = int(max(no.value, nc.value)) + 16
                continue
            self._check(rc)
            return np.sort(opens[:no.value]), np.sort(closes[:nc.value])

    def text_scan_chunks(self, text, slot: int = 0, n_bytes: Optional[int] = None, mem_kind: int = ffi.GPR_MEM_HOST):
        """generator over the pipelined scan: yields (opens, closes, bytes_done) per 4 MB chunk while later chunks are
        still being uploaded (gpr_text_scan_begin / gpr_text_scan_next)"""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
        n = int(buf.size if n_bytes is None else n_bytes)
        self._check(self._lib.gpr_text_scan_begin(self._h, slot, _ptr(buf), n, mem_kind))
        cap = 1 << 14
        opens, closes = np.empty(cap, np.uint64), np.empty(cap, np.uint64)
        no, nc, done, more = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int32(1)
        while more.value:
            self._check(self._lib.gpr_text_scan_next(self._h, _ptr(opens), _ptr(closes), cap, C.byref(no), C.byref(nc),
                                                     C.byref(done), C.byref(more)))
            yield opens[:no.value].copy(), closes[:nc.value].copy(), done.value

    def text_parse(self, spans: np.ndarray, t_end: int, step: int, T: int, n_rows: int, slot: int = 0,
                   plane: int = 0, fill: bool = True, window_seconds: Optional[int] = None,
                   resident: bool = False) -> np.ndarray:
        """Parse the samples of ``spans`` (structured array of SPAN_DTYPE, sorted by begin) of the text in
        ``slot`` into the context's plane (or, ``resident=True``, the resident ring); returns the spans with
        their out-fields filled.  ``window_seconds`` defaults to ``T * step``."""
        spans = np.ascontiguousarray(spans, dtype=self.SPAN_DTYPE)
        g = ffi.gpr_text_grid()
        g.struct_size = C.sizeof(ffi.gpr_text_grid)
        g.flags = (ffi.GPR_TEXT_FILL if fill and not resident else 0) | (ffi.GPR_TEXT_RESIDENT if resident else 0)
        g.t_end, g.step = int(t_end), int(step)
        g.window_seconds = int(T) * int(step) if window_seconds is None else int(window_seconds)
        g.n_samples, g.n_rows = int(T), int(n_rows)
        self._check(self._lib.gpr_text_parse(self._h, slot, _ptr(spans), len(spans), C.byref(g), plane))
        return spans

    def resident_head(self) -> int:
        h = C.c_uint32()
        self._check(self._lib.gpr_resident_head(self._h, C.byref(h)))
        return h.value

    def resident_advance(self, n_new: int):
        """open the next ``n_new`` buckets of the resident ring without data (all rows: no sample)"""
        self._check(self._lib.gpr_resident_advance(self._h, int(n_new)))

    def text_planes(self):
        u, w = C.c_void_p(), C.c_void_p()
        self._check(self._lib.gpr_text_planes(self._h, C.byref(u), C.byref(w)))
        return u.value, w.value

    def synth_fill(self, seed: int, plane: int, dst, pod_offset: int, P: int, G: int, T: int,
                   row_stride: int = 0):
        self._check(self._lib.gpr_synth_fill(self._h, seed, plane, _ptr(dst), pod_offset, P, G, T,
                                             row_stride))

    def synth_eligible(self, seed: int, dst, pod_offset: int, P: int):
        self._check(self._lib.gpr_synth_eligible(self._h, seed, _ptr(dst), pod_offset, P))
