"""ctypes binding of lib3dinfomax_b200.so (C ABI declared in include/i3d.h) + in-tree nvcc build.

There is NO CPU fallback: if the shared library is missing or a symbol fails to resolve, importing
the compute path raises.  Every wrapper takes torch tensors, checks device / dtype / contiguity,
passes raw device pointers plus the caller's current CUDA stream, and turns a non-zero return code
into ``RuntimeError(i3d_last_error_string())``.
"""
import ctypes
import os
import re
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(_ROOT, "include", "i3d.h")
SO_PATH = os.path.join(_HERE, "lib3dinfomax_b200.so")
SOURCES = ["i3d_runtime.cu", "i3d_graph.cu", "i3d_pna.cu", "i3d_bn.cu", "i3d_bond.cu", "i3d_net3d.cu", "i3d_gemm.cu", "i3d_gemm_tc.cu", "i3d_gemm_tc_ws.cu", "i3d_gemm_tc_tn.cu",
           "i3d_loss.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "0"]

_lock = threading.Lock()
_lib = None


class gemm_seg(ctypes.Structure):
    """mirror of i3d_gemm_seg (include/i3d.h)"""
    _fields_ = [("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("a_idx", ctypes.c_void_p),
                ("b_idx", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("K", ctypes.c_int32),
                ("lda", ctypes.c_int32), ("ldb", ctypes.c_int32)]


class reduce_ws(ctypes.Structure):
    """struct i3d_reduce_ws of include/i3d.h"""
    _fields_ = [("slots", ctypes.c_void_p), ("slot_bytes", ctypes.c_int64), ("counter", ctypes.c_void_p)]


class prep_item(ctypes.Structure):
    """struct i3d_prep_item of include/i3d.h"""
    _fields_ = [("B", ctypes.c_void_p), ("hi", ctypes.c_void_p), ("lo", ctypes.c_void_p), ("ldb", ctypes.c_int32),
                ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("kpad", ctypes.c_int32), ("ldo", ctypes.c_int32),
                ("col0", ctypes.c_int32), ("transposed", ctypes.c_int32), ("tile0", ctypes.c_int32)]


_P, _I, _L, _F, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); must cover every function declared in include/i3d.h (tests check this)
SIGNATURES = {
    "i3d_version": (_I, []),
    "i3d_last_error_string": (ctypes.c_char_p, []),
    "i3d_launch_count": (_L, []),
    "i3d_set_pdl": (_I, [_I]),
    "i3d_csr_build": (_I, [_P, _P, _L, _L, _P, _P, _P, _P, _P, _P]),
    "i3d_csr_build_i32": (_I, [_P, _P, _L, _L, _P, _P, _P, _P, _P, _P]),
    "i3d_segment_ptr": (_I, [_P, _L, _P, _P]),
    "i3d_degree_scalers": (_I, [_P, _L, _P, _P, _P]),
    "i3d_degree_scalers_avg": (_I, [_P, _L, _D, _P, _P, _P]),
    "i3d_scale_rows": (_I, [_P, _I, _P, _L, _I, _P, _I, _P]),
    "i3d_degree_plan": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P]),
    "i3d_posttrans_merge": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P]),
    "i3d_gemm_nt_bucketed": (_I, [_L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _I,
                                  _P]),
    "i3d_gemm_tn_chunked": (_I, [_L, _I, ctypes.POINTER(gemm_seg), _P, _I, _L, _P, _I, _P]),
    "i3d_gemm_debug_counters": (_I, [_P]),
    "i3d_posttrans_unmerge": (_I, [_P, _I, _I, _I, _P, _I, _P]),
    "i3d_collate_2d": (_I, [_P, _L, _P, _P, _P, _L, _P, _I, _P, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
    "i3d_collate_3d": (_I, [_P, _L, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    "i3d_collate_2d_struct": (_I, [_P, _L, _P, _P, _P, _L, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P, _P,
                                   _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "i3d_collate_3d_struct": (_I, [_P, _L, _I, _P, _P, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "i3d_gemm_ex_v": (_I, [_I, _L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _I, _P, ctypes.c_size_t, _P, _I, _P,
                           _P]),
    "i3d_gemm_nt_prepared_v": (_I, [_L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _I, _P, _P, _I, _P, _P]),
    "i3d_gemm_nt_bucketed_v": (_I, [_L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _I,
                                    _P, _P]),
    "i3d_act_colstats_v": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "i3d_bn_apply_v": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _I, _P, _P]),
    "i3d_bn_bwd_reduce_v": (_I, [_P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _I, _P, _P, _P]),
    "i3d_bn_bwd_fused_v": (_I, [_P, _I, _P, _I, _L, _I, _I, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _P]),
    "i3d_bn_bwd_apply_v": (_I, [_P, _I, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P]),
    "i3d_embed_sum_fwd": (_I, [_P, _L, _I, _P, _P, _P, _I, _P, _P]),
    "i3d_embed_sum_bwd": (_I, [_P, _L, _I, _P, _P, _P, _I, _P, _I, _I, _P]),
    "i3d_gemm_backend": (_I, [_I]),
    "i3d_transpose": (_I, [_P, _L, _I, _I, _P, _I, _P]),
    "i3d_gemm": (_I, [_I, _L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _I, _P]),
    "i3d_gemm_ws_bytes": (ctypes.c_size_t, [_I, _L, _I, _I, ctypes.POINTER(gemm_seg)]),
    "i3d_gemm_ex": (_I, [_I, _L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _I, _P, ctypes.c_size_t, _P, _I, _P]),
    "i3d_gemm_prep_describe": (_I, [_I, _I, ctypes.POINTER(gemm_seg), _I, _P, _I, ctypes.POINTER(prep_item),
                                    ctypes.POINTER(ctypes.c_int)]),
    "i3d_gemm_prep_run": (_I, [_P, _I, _I, _P]),
    "i3d_gemm_nt_prepared_ok": (_I, [_L, _I, _I, ctypes.POINTER(gemm_seg)]),
    "i3d_gemm_nt_prepared": (_I, [_L, _I, _I, ctypes.POINTER(gemm_seg), _P, _I, _P, _I, _P, _P, _I, _P]),
    "i3d_act_colstats": (_I, [_P, _L, _I, _I, _I, _P, _P]),
    "i3d_bn_apply": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _I, _P]),
    "i3d_bn_bwd_reduce": (_I, [_P, _I, _P, _I, _L, _I, _I, _P, _P, _P]),
    "i3d_bn_bwd_reduce_ex": (_I, [_P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _I, _P]),
    "i3d_bn_bwd_apply": (_I, [_P, _I, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P]),
    "i3d_edge_gather_add": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _L, _I, _P, _I, _P, _I, _P, _P, _P]),
    "i3d_bond_tables_fwd": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    "i3d_bond_tables_bwd": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "i3d_act_fwd": (_I, [_P, _L, _I, _P, _P]),
    "i3d_act_bwd": (_I, [_P, _P, _L, _I, _P, _P]),
    "i3d_pna_aggregate_fwd": (_I, [_P, _P, _L, _I, _P, _I, _P]),
    "i3d_pna_aggregate_bwd": (_I, [_P, _I, _P, _P, _I, _P, _L, _I, _P, _P]),
    "i3d_segment_readout_fwd": (_I, [_P, _I, _P, _L, _I, _I, _P, _P, _P]),
    "i3d_segment_readout_bwd": (_I, [_P, _P, _I, _P, _P, _L, _I, _I, _P, _P, _I, _P]),
    "i3d_segment_readout_bwd_v": (_I, [_P, _P, _I, _P, _P, _L, _I, _I, _P, _P, _I, _L, _P]),
    "i3d_segment_sum_fwd": (_I, [_P, _I, _P, _P, _L, _I, _I, _P, _I, _P, _I, _P]),
    "i3d_segment_sum_bwd": (_I, [_P, _P, _P, _L, _I, _I, _P, _P]),
    "i3d_fourier_encode": (_I, [_P, _P, _L, _I, _P, _P]),
    "i3d_soft_gate_fwd": (_I, [_P, _L, _I, _P, _P, _P, _P, _P]),
    "i3d_soft_gate_bwd": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _P]),
    "i3d_broadcast_rows": (_I, [_P, _L, _I, _P, _P]),
    "i3d_colsum": (_I, [_P, _I, _L, _I, _P, _P]),
    "i3d_add": (_I, [_P, _P, _L, _P, _P]),
    "i3d_add_rows": (_I, [_P, _I, _P, _I, _L, _I, _P, _I, _P]),
    "i3d_row_norms": (_I, [_P, _L, _I, _P, _P]),
    "i3d_ntxent_rows_fwd": (_I, [_P, _L, _L, _I, _P, _P, _I, _F, _F, _L, _P, _P, _P]),
    "i3d_contrastive_metrics": (_I, [_P, _L, _P, _P, _F, _P, _P, _P]),
    "i3d_embedding_metrics": (_I, [_P, _L, _P, _L, _I, _F, _F, _P, _P, _P]),
    "i3d_sum_scaled": (_I, [_P, _L, _F, _P, _P]),
    "i3d_ntxent_rows_bwd": (_I, [_P, _L, _L, _I, _P, _P, _I, _F, _F, _L, _P, _P, _F, _P, _P, _P]),
    "i3d_norm_bwd_accum": (_I, [_P, _P, _P, _L, _I, _P, _P]),
    "i3d_adam_step": (_I, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _D, _D, _L, _P, _P, _P]),
    "i3d_adam_step_nvls": (_I, [_P, _P, _L, _I, _I, _D, _D, _D, _D, _D, _D, _L, _P, _P, _P]),
    "i3d_add_i64": (_I, [_P, _L, _P]),
    "i3d_multi_copy": (_I, [_P, _P, _P, _I, _P, _I, _P]),
    "i3d_multi_copy_strided": (_I, [_P, _P, _P, _P, _I, _P, _I, _P]),
}


def declared_symbols():
    """Function names declared in include/i3d.h."""
    with open(HEADER) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(i3d_[a-z0-9_]+)\s*\(", text)))


def _needs_build():
    if not os.path.isfile(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into 3dinfomax_b200/lib3dinfomax_b200.so (nvcc cross-compiles without a GPU)."""
    with _lock:
        if not force and not _needs_build():
            return SO_PATH
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("I3D_NVCC_EXTRA", "").split()          # tuning builds, e.g. -DI3D_WS_PREFETCH=6
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-shared", "-o", SO_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout)
        return SO_PATH


def load():
    """dlopen the library (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if _needs_build():
        try:
            build()
        except FileNotFoundError as e:  # no nvcc on this box and no prebuilt .so
            if not os.path.isfile(SO_PATH):
                raise RuntimeError("lib3dinfomax_b200.so is missing and nvcc is unavailable: the 3dinfomax_b200 compute "
                                   "path has no CPU fallback") from e
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().i3d_last_error_string().decode("utf-8", "replace")


def launch_count():
    return int(load().i3d_launch_count())


def check(rc, what):
    if rc != 0:
        msg = last_error()
        if "Expected more than 1 value per channel" in msg:
            raise ValueError(msg)
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))
