"""petsc-openacc_b200 -- Blackwell-native SeqAIJ sparse mat-vec hot path.

This Python module is only a ctypes view of the C ABI declared in include/b200_seqaij.h
(libb200aij.so, hand-written CUDA for sm_100a) for tests and bench.py.  The product is the
shared library and the PETSc-named C symbols on top of it (host/); PyTorch is used by callers for
device memory, streams and torch.distributed only.

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every
compute entry point returns an error on a machine without an sm_100 device, and a missing
library raises at import time.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libb200aij.so")

MODE_FAST, MODE_EXACT, MODE_EXACT_FMA = 0, 1, 2
KERNEL_AUTO, KERNEL_ROW, KERNEL_STREAM, KERNEL_VECTOR, KERNEL_MERGE, KERNEL_CPROW, KERNEL_SELL = range(7)
KERNEL_NAMES = {0: "auto", 1: "row", 2: "stream", 3: "vector", 4: "merge", 5: "cprow", 6: "sell"}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200 error {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile libb200aij.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-s"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: run `make -C petsc-openacc_b200/csrc` (or "
        "__graft_entry__.build()). There is no Python/CPU fallback for the CUDA path.")

lib = C.CDLL(LIB_PATH)


class CsrInfo(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("nz", C.c_int32),
                ("nonzerorowcnt", C.c_int32), ("rmax", C.c_int32),
                ("compressedrow_use", C.c_int32), ("cprow_nrows", C.c_int32),
                ("kernel_fast", C.c_int32), ("kernel_exact", C.c_int32),
                ("vector_lanes", C.c_int32), ("stream_tiles", C.c_int32),
                ("merge_tiles", C.c_int32), ("has_transpose", C.c_int32),
                ("index8_diagonals", C.c_int32),
                ("hist", C.c_int32 * 16), ("device_bytes", C.c_uint64),
                ("sell_chunks", C.c_int32), ("sell_sigma", C.c_int32), ("sell_padded_nnz", C.c_uint64),
                ("rowlen8", C.c_int32), ("reserved", C.c_int32)]


class CgResult(C.Structure):
    _fields_ = [("its", C.c_int32), ("reason", C.c_int32), ("rnorm", C.c_double),
                ("rnorm0", C.c_double), ("solve_ms", C.c_double), ("launches", C.c_uint64)]


lib.b200_last_error.restype = C.c_char_p
lib.b200_version.restype = C.c_char_p
lib.b200_launch_count.restype = C.c_uint64

# every symbol include/b200_seqaij.h declares; tests check they are all exported
ABI_SYMBOLS = [
    "b200_init", "b200_last_error", "b200_launch_count", "b200_device_sm_count", "b200_version",
    "b200_csr_create", "b200_csr_create_from_device", "b200_csr_update_values",
    "b200_csr_destroy", "b200_csr_get_info", "b200_csr_set_kernel", "b200_csr_build_transpose",
    "b200_csr_device_arrays", "b200_spmv", "b200_spmv_add", "b200_spmv_transpose",
    "b200_spmv_transpose_add", "b200_spmv_residual", "b200_spmv_jacobi_sweep", "b200_spmv_host", "b200_spmv_add_host",
    "b200_spmv_transpose_host", "b200_spmv_transpose_add_host", "b200_host_alloc", "b200_host_free", "b200_host_register",
    "b200_host_unregister", "b200_vec_set", "b200_vec_copy", "b200_vec_axpy", "b200_vec_aypx",
    "b200_vec_pointwise_mult", "b200_vec_dot", "b200_vec_norm2", "b200_vec_norm_inf",
    "b200_vec_sum", "b200_cg_jacobi", "b200_gen_vector",
    "b200_csr_build_sell", "b200_sell_pack_size", "b200_sell_pack",
    "b200_wmerge_plan_size", "b200_wmerge_plan", "b200_colblock_split",
]


def check(rc):
    if rc != 0:
        raise B200Error(rc, (lib.b200_last_error() or b"").decode())


def launch_count():
    return int(lib.b200_launch_count())


def init(device=0):
    check(lib.b200_init(C.c_int(device)))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _dptr(t):
    """device pointer of a torch CUDA tensor (float64/int32, contiguous) or a raw int"""
    if isinstance(t, int):
        return C.c_void_p(t)
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _stream(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)


class Csr:
    """Device-resident mirror of a SeqAIJ matrix (a->i, a->j, a->a) plus its kernel plan."""

    def __init__(self, ai, aj, aa, n=None):
        ai = np.ascontiguousarray(ai, dtype=np.int32)
        aj = np.ascontiguousarray(aj, dtype=np.int32)
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        self.m = len(ai) - 1
        self.n = self.m if n is None else int(n)
        self.nz = int(ai[-1])
        self._h = C.c_void_p(0)
        check(lib.b200_csr_create(C.byref(self._h), C.c_int32(self.m), C.c_int32(self.n),
                                  _np_ptr(ai), _np_ptr(aj), _np_ptr(aa)))

    @classmethod
    def from_device(cls, d_ai, d_aj, d_aa, m, n):
        self = cls.__new__(cls)
        self.m, self.n = int(m), int(n)
        self._h = C.c_void_p(0)
        check(lib.b200_csr_create_from_device(C.byref(self._h), C.c_int32(m), C.c_int32(n),
                                              _dptr(d_ai), _dptr(d_aj), _dptr(d_aa)))
        self.nz = self.info().nz
        return self

    def destroy(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.b200_csr_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def info(self):
        i = CsrInfo()
        check(lib.b200_csr_get_info(self._h, C.byref(i)))
        return i

    def set_kernel(self, kernel):
        check(lib.b200_csr_set_kernel(self._h, C.c_int(kernel)))

    def update_values(self, aa):
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        check(lib.b200_csr_update_values(self._h, _np_ptr(aa)))

    def build_transpose(self):
        check(lib.b200_csr_build_transpose(self._h))

    def build_sell(self, sigma=1):
        """The optional SELL-32-sigma copy; use it with set_kernel(KERNEL_SELL)."""
        check(lib.b200_csr_build_sell(self._h, C.c_int32(sigma)))

    # device-resident vectors (torch CUDA tensors)
    def mult(self, x, y, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv(self._h, _dptr(x), _dptr(y), C.c_int(mode), _stream(stream)))
        return y

    def mult_add(self, x, y, z, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv_add(self._h, _dptr(x), _dptr(y), _dptr(z), C.c_int(mode),
                                _stream(stream)))
        return z

    def mult_transpose(self, x, y, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv_transpose(self._h, _dptr(x), _dptr(y), C.c_int(mode),
                                      _stream(stream)))
        return y

    def mult_transpose_add(self, x, z, y, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv_transpose_add(self._h, _dptr(x), _dptr(z), _dptr(y), C.c_int(mode),
                                          _stream(stream)))
        return y

    def residual(self, x, b, r, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv_residual(self._h, _dptr(x), _dptr(b), _dptr(r), C.c_int(mode), _stream(stream)))
        return r

    def jacobi_sweep(self, x, b, dinv, xnew, mode=MODE_FAST, stream=None):
        check(lib.b200_spmv_jacobi_sweep(self._h, _dptr(x), _dptr(b), _dptr(dinv), _dptr(xnew), C.c_int(mode),
                                         _stream(stream)))
        return xnew

    # host vectors (numpy): the MatMult_SeqAIJ(Mat,Vec,Vec) shape, synchronous
    def mult_host(self, x, y=None, mode=MODE_FAST):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty(self.m, dtype=np.float64)
        check(lib.b200_spmv_host(self._h, _np_ptr(x), _np_ptr(y), C.c_int(mode)))
        return y

    def mult_add_host(self, x, y, z=None, mode=MODE_FAST):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if z is None:
            z = np.empty(self.m, dtype=np.float64)
        check(lib.b200_spmv_add_host(self._h, _np_ptr(x), _np_ptr(y), _np_ptr(z), C.c_int(mode)))
        return z

    def mult_transpose_host(self, x, y=None, mode=MODE_FAST):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty(self.n, dtype=np.float64)
        check(lib.b200_spmv_transpose_host(self._h, _np_ptr(x), _np_ptr(y), C.c_int(mode)))
        return y

    def cg_jacobi(self, b, x, rtol=1e-14, atol=1e-12, max_it=10000, mode=MODE_FAST, stream=None):
        res = CgResult()
        check(lib.b200_cg_jacobi(self._h, _dptr(b), _dptr(x), C.c_double(rtol), C.c_double(atol),
                                 C.c_int32(max_it), C.c_int(mode), C.byref(res), _stream(stream)))
        return res


def sell_pack(ai, aj, aa, sigma=1):
    """Host packing of the SELL-32-sigma copy (no device needed): (cs, perm, val, col)."""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    m = len(ai) - 1
    nchunks, padded = C.c_int32(0), C.c_uint64(0)
    check(lib.b200_sell_pack_size(C.c_int32(m), _np_ptr(ai), C.c_int32(sigma), C.byref(nchunks), C.byref(padded)))
    cs = np.zeros(nchunks.value + 1, dtype=np.uint32)
    perm = np.zeros(max(nchunks.value * 32, 1), dtype=np.int32)
    val = np.zeros(max(padded.value, 1), dtype=np.float64)
    col = np.zeros(max(padded.value, 1), dtype=np.int32)
    check(lib.b200_sell_pack(C.c_int32(m), _np_ptr(ai), _np_ptr(aj), _np_ptr(aa), C.c_int32(sigma), _np_ptr(cs), _np_ptr(perm),
                             _np_ptr(val), _np_ptr(col)))
    return cs, perm[:nchunks.value * 32], val[:padded.value], col[:padded.value]


def wmerge_plan(ai):
    """Host-only plan of k_wmerge: (chunks[n, 4], blk[nblocks + 1])."""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    m = len(ai) - 1
    nc, nb = C.c_int32(0), C.c_int32(0)
    check(lib.b200_wmerge_plan_size(C.c_int32(m), _np_ptr(ai), C.byref(nc), C.byref(nb)))
    chunks = np.zeros((max(nc.value, 1), 4), np.int32)
    blk = np.zeros(nb.value + 1, np.int32)
    check(lib.b200_wmerge_plan(C.c_int32(m), _np_ptr(ai), _np_ptr(chunks), _np_ptr(blk)))
    return chunks[:nc.value], blk


def colblock_split(ai, aj, n, limit_bytes):
    """Host-only column-block split of the skewed plan: (nblocks, split[nblocks, m])."""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    m = len(ai) - 1
    nb = C.c_int32(0)
    check(lib.b200_colblock_split(C.c_int32(m), C.c_int32(n), _np_ptr(ai), _np_ptr(aj), C.c_int64(limit_bytes), C.byref(nb), None))
    split = np.zeros((max(nb.value, 1), max(m, 1)), np.int32)
    if nb.value:
        check(lib.b200_colblock_split(C.c_int32(m), C.c_int32(n), _np_ptr(ai), _np_ptr(aj), C.c_int64(limit_bytes), C.byref(nb), _np_ptr(split)))
    return nb.value, split[:nb.value, :m]


def gen_vector(n, seed=0xB200):
    x = np.empty(n, dtype=np.float64)
    check(lib.b200_gen_vector(_np_ptr(x), C.c_int64(n), C.c_uint64(seed)))
    return x


class PinnedArray:
    """float64 numpy view of page-locked host memory from b200_host_alloc."""

    def __init__(self, n):
        self._p = C.c_void_p(0)
        self.n = int(n)
        check(lib.b200_host_alloc(C.byref(self._p), C.c_size_t(max(self.n, 1) * 8)))
        buf = (C.c_double * self.n).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=self.n)

    def free(self):
        if self._p.value:
            self.array = None
            lib.b200_host_free(self._p)
            self._p = C.c_void_p(0)


# vector ops on torch CUDA tensors; reductions write into a 1-element CUDA tensor
def vec_dot(x, y, out, stream=None):
    check(lib.b200_vec_dot(_dptr(x), _dptr(y), C.c_int64(x.numel()), _dptr(out), _stream(stream)))


def vec_norm2(x, out, stream=None):
    check(lib.b200_vec_norm2(_dptr(x), C.c_int64(x.numel()), _dptr(out), _stream(stream)))


def vec_norm_inf(x, out, stream=None):
    check(lib.b200_vec_norm_inf(_dptr(x), C.c_int64(x.numel()), _dptr(out), _stream(stream)))


def vec_sum(x, out, stream=None):
    check(lib.b200_vec_sum(_dptr(x), C.c_int64(x.numel()), _dptr(out), _stream(stream)))


def vec_axpy(y, a, x, stream=None):
    check(lib.b200_vec_axpy(_dptr(y), C.c_double(a), _dptr(x), C.c_int64(x.numel()),
                            _stream(stream)))


def vec_aypx(y, a, x, stream=None):
    check(lib.b200_vec_aypx(_dptr(y), C.c_double(a), _dptr(x), C.c_int64(x.numel()),
                            _stream(stream)))


def vec_set(x, a, stream=None):
    check(lib.b200_vec_set(_dptr(x), C.c_double(a), C.c_int64(x.numel()), _stream(stream)))


def vec_copy(y, x, stream=None):
    check(lib.b200_vec_copy(_dptr(y), _dptr(x), C.c_int64(x.numel()), _stream(stream)))


def vec_pointwise_mult(w, x, y, stream=None):
    check(lib.b200_vec_pointwise_mult(_dptr(w), _dptr(x), _dptr(y), C.c_int64(x.numel()),
                                      _stream(stream)))


# ---- MatMult_MPIAIJ (include/b200_mpiaij.h) ----------------------------------------------------
MPIAIJ_SYMBOLS = [
    "b200_mpiaij_create", "b200_mpiaij_destroy", "b200_mpiaij_get_sizes", "b200_mpiaij_get_garray",
    "b200_mpiaij_get_recv_offsets", "b200_mpiaij_copy_block", "b200_mpiaij_set_peer_garray",
    "b200_mpiaij_get_send_list", "b200_mpiaij_upload", "b200_mpiaij_get_blocks",
    "b200_mpiaij_window_ipc_handle", "b200_mpiaij_window_ptr", "b200_mpiaij_open_peer_window",
    "b200_mpiaij_set_peer_window", "b200_mpiaij_mult_begin", "b200_mpiaij_mult_local",
    "b200_mpiaij_mult_end", "b200_mpiaij_mult", "b200_mpiaij_pack", "b200_mpiaij_mult_add_ghost",
    "b200_mpiaij_check", "b200_mpiaij_mult_host", "b200_mpiaij_mult_finish",
    "b200_mpiaij_set_rank_window", "b200_mpiaij_allreduce_sum", "b200_mpiaij_cg_jacobi",
    "b200_mpiaij_pattern_symmetric", "b200_mpiaij_tile_schedule",
]
ABI_SYMBOLS += MPIAIJ_SYMBOLS


class MpiAij:
    """One rank's Mat_MPIAIJ {A, B, garray, lvec}: host split + (after upload) device blocks."""

    def __init__(self, size, rank, base, ai, aj_global, aa):
        self.size, self.rank = int(size), int(rank)
        self.base = np.ascontiguousarray(base, dtype=np.int32)
        ai = np.ascontiguousarray(ai, dtype=np.int32)
        aj = np.ascontiguousarray(aj_global, dtype=np.int32)
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        self._h = C.c_void_p(0)
        check(lib.b200_mpiaij_create(C.byref(self._h), C.c_int32(size), C.c_int32(rank),
                                     _np_ptr(self.base), _np_ptr(ai), _np_ptr(aj), _np_ptr(aa)))
        s = np.zeros(6, np.int32)
        check(lib.b200_mpiaij_get_sizes(self._h, _np_ptr(s)))
        self.nloc, self.annz, self.bnnz, self.nghost, self.brows, self.nsrc = (int(v) for v in s)

    def destroy(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.b200_mpiaij_destroy(self._h)
            self._h = C.c_void_p(0)

    def garray(self):
        g = np.zeros(max(self.nghost, 1), np.int32)
        check(lib.b200_mpiaij_get_garray(self._h, _np_ptr(g)))
        return g[:self.nghost]

    def recv_offsets(self):
        off = np.zeros(self.size + 1, np.int32)
        check(lib.b200_mpiaij_get_recv_offsets(self._h, _np_ptr(off)))
        return off

    def block(self, which):
        nnz = self.bnnz if which else self.annz
        ai = np.zeros(self.nloc + 1, np.int32)
        aj = np.zeros(max(nnz, 1), np.int32)
        aa = np.zeros(max(nnz, 1))
        check(lib.b200_mpiaij_copy_block(self._h, C.c_int(which), _np_ptr(ai), _np_ptr(aj), _np_ptr(aa)))
        return ai, aj[:nnz], aa[:nnz]

    def set_peer_garray(self, peer, garray):
        g = np.ascontiguousarray(garray, dtype=np.int32)
        check(lib.b200_mpiaij_set_peer_garray(self._h, C.c_int32(peer), _np_ptr(g), C.c_int32(len(g))))

    def send_list(self, peer):
        cnt, off = C.c_int32(0), C.c_int32(0)
        check(lib.b200_mpiaij_get_send_list(self._h, C.c_int32(peer), C.byref(cnt), None, C.byref(off)))
        idx = np.zeros(max(cnt.value, 1), np.int32)
        check(lib.b200_mpiaij_get_send_list(self._h, C.c_int32(peer), C.byref(cnt), _np_ptr(idx), C.byref(off)))
        return idx[:cnt.value], off.value

    def upload(self):
        check(lib.b200_mpiaij_upload(self._h))

    def ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        check(lib.b200_mpiaij_window_ipc_handle(self._h, buf))
        return bytes(buf)

    def window_ptr(self):
        p = C.c_void_p(0)
        check(lib.b200_mpiaij_window_ptr(self._h, C.byref(p)))
        return p.value

    def open_peer_window(self, peer, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        check(lib.b200_mpiaij_open_peer_window(self._h, C.c_int32(peer), buf))

    def set_peer_window(self, peer, ptr):
        check(lib.b200_mpiaij_set_peer_window(self._h, C.c_int32(peer), C.c_void_p(ptr)))

    def mult_begin(self, x, stream=None):
        check(lib.b200_mpiaij_mult_begin(self._h, _dptr(x), _stream(stream)))

    def mult_local(self, x, y, mode=MODE_FAST, stream=None):
        check(lib.b200_mpiaij_mult_local(self._h, _dptr(x), _dptr(y), C.c_int(mode), _stream(stream)))

    def mult_end(self, y, mode=MODE_FAST, stream=None):
        check(lib.b200_mpiaij_mult_end(self._h, _dptr(y), C.c_int(mode), _stream(stream)))

    def mult(self, x, y, mode=MODE_FAST, stream=None):
        check(lib.b200_mpiaij_mult(self._h, _dptr(x), _dptr(y), C.c_int(mode), _stream(stream)))

    def mult_finish(self, x, y, mode=MODE_FAST, stream=None):
        check(lib.b200_mpiaij_mult_finish(self._h, _dptr(x), _dptr(y), C.c_int(mode), _stream(stream)))

    def mult_host(self, hx, hy, mode=MODE_FAST):
        check(lib.b200_mpiaij_mult_host(self._h, _np_ptr(hx), _np_ptr(hy), C.c_int(mode)))

    def set_rank_window(self, q, handle=None, ptr=None):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        check(lib.b200_mpiaij_set_rank_window(self._h, C.c_int32(q), buf, C.c_void_p(ptr or 0)))

    def allreduce_sum(self, vals, stream=None):
        check(lib.b200_mpiaij_allreduce_sum(self._h, _dptr(vals), C.c_int32(vals.numel()), _stream(stream)))

    def cg_jacobi(self, b, x, rtol=1e-14, atol=1e-12, max_it=10000, mode=MODE_FAST, stream=None):
        res = CgResult()
        check(lib.b200_mpiaij_cg_jacobi(self._h, _dptr(b), _dptr(x), C.c_double(rtol), C.c_double(atol),
                                        C.c_int32(max_it), C.c_int(mode), C.byref(res), _stream(stream)))
        return res

    def pack(self, peer, x, buf, stream=None):
        check(lib.b200_mpiaij_pack(self._h, C.c_int32(peer), _dptr(x), _dptr(buf), _stream(stream)))

    def mult_add_ghost(self, lvec, y, mode=MODE_FAST, stream=None):
        check(lib.b200_mpiaij_mult_add_ghost(self._h, _dptr(lvec), _dptr(y), C.c_int(mode), _stream(stream)))

    def check(self):
        check(lib.b200_mpiaij_check(self._h))

    def pattern_symmetric(self):
        """(True, -1) or (False, first peer this rank sends to but does not receive from)."""
        peer = C.c_int32(-1)
        rc = lib.b200_mpiaij_pattern_symmetric(self._h, C.byref(peer))
        return rc == 0, peer.value


def mpiaij_tile_schedule(ghost_rows_per_tile, grid, npush=0, push_charge=6.0):
    """Host-only tile -> CTA schedule of the fused MatMult_MPIAIJ launch."""
    g = np.ascontiguousarray(ghost_rows_per_tile, dtype=np.int32)
    out = np.zeros(max(len(g), 1), np.int32)
    check(lib.b200_mpiaij_tile_schedule(C.c_int32(len(g)), _np_ptr(g), C.c_int32(grid), C.c_int32(npush),
                                        C.c_double(push_charge), _np_ptr(out)))
    return out[:len(g)]


def gen_poisson7(M, size=1, rank=0, refpoint=True, vectors=False):
    """This rank's rows of the reference problem (src/helper.cpp) from the product generator."""
    out = np.zeros(12, np.int32)
    check(lib.b200_gen_poisson7_info(M, M, M, size, rank, _np_ptr(out)))
    nloc, rstart, nnz = int(out[9]), int(out[10]), int(out[11])
    ai, aj, aa = np.zeros(nloc + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
    rhs = np.zeros(nloc) if vectors else None
    ex = np.zeros(nloc) if vectors else None
    check(lib.b200_gen_poisson7(M, M, M, size, rank, int(refpoint), _np_ptr(ai), _np_ptr(aj),
                                _np_ptr(aa), _np_ptr(rhs) if vectors else None,
                                _np_ptr(ex) if vectors else None))
    base = np.zeros(size + 1, np.int32)
    check(lib.b200_gen_poisson7_bases(M, M, M, size, _np_ptr(base)))
    return dict(ai=ai, aj=aj, aa=aa, rhs=rhs, exact=ex, rstart=rstart, base=base, info=out)


ABI_SYMBOLS += ["b200_gen_poisson7_info", "b200_gen_poisson7_bases", "b200_gen_poisson7",
                "b200_gen_powerlaw_rowptr", "b200_gen_powerlaw_fill", "b200_gen_stencil27"]


def gen_stencil27(N, seed=0):
    """The 27-point matrix of BASELINE configs[3] from the product generator (threaded C++)."""
    n = N ** 3
    ai = np.zeros(n + 1, np.int32)
    check(lib.b200_gen_stencil27(C.c_int32(N), C.c_uint64(seed), _np_ptr(ai), None, None))
    aj, aa = np.zeros(int(ai[-1]), np.int32), np.zeros(int(ai[-1]))
    check(lib.b200_gen_stencil27(C.c_int32(N), C.c_uint64(seed), _np_ptr(ai), _np_ptr(aj), _np_ptr(aa)))
    return ai, aj, aa


def gen_powerlaw(m, n=None, alpha=2.0, lmax=10000, seed=0x5EED):
    """The irregular matrix of BASELINE configs[4] from the product generator (threaded C++)."""
    n = m if n is None else n
    ai = np.zeros(m + 1, np.int32)
    check(lib.b200_gen_powerlaw_rowptr(C.c_int32(m), C.c_int32(n), C.c_double(alpha), C.c_int32(lmax),
                                       C.c_uint64(seed), _np_ptr(ai)))
    aj = np.zeros(int(ai[-1]), np.int32)
    aa = np.zeros(int(ai[-1]))
    check(lib.b200_gen_powerlaw_fill(C.c_int32(m), C.c_int32(n), C.c_double(alpha), C.c_int32(lmax),
                                     C.c_uint64(seed), _np_ptr(ai), _np_ptr(aj), _np_ptr(aa)))
    return ai, aj, aa
