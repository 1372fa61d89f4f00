"""ctypes binding of libesr_b200.so (include/esr_b200.h).  The product path has no fallback: if the
shared library is missing, or the device is not sm_100, every op raises."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "libesr_b200.so")
CSRC = os.path.join(PKG_ROOT, "csrc")

ESR_F16, ESR_BF16, ESR_BF16X3 = 0, 1, 2


class EsrError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    """mirror of esr_conv3x3_args"""
    _fields_ = [
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("dtype", C.c_int),
        ("in_", C.c_void_p), ("in_planes_total", C.c_int), ("in_plane_off", C.c_int), ("cin_planes", C.c_int),
        ("wpacked", C.c_void_p), ("bias", C.c_void_p), ("cout", C.c_int), ("cout_pad", C.c_int), ("kcp", C.c_int),
        ("lrelu", C.c_int), ("slope", C.c_float), ("alpha", C.c_float),
        ("res1", C.c_void_p), ("res1_is16", C.c_int), ("res1_planes_total", C.c_int), ("res1_plane_off", C.c_int), ("beta1", C.c_float),
        ("res2", C.c_void_p), ("res2_planes_total", C.c_int), ("res2_plane_off", C.c_int), ("beta2", C.c_float),
        ("res3", C.c_void_p), ("res3_planes_total", C.c_int), ("res3_plane_off", C.c_int), ("beta3", C.c_float),
        ("lead_planes", C.c_int), ("lead_acc", C.c_void_p), ("lead_planes_total", C.c_int),
        ("mask16", C.c_void_p), ("mask_planes_total", C.c_int), ("mask_plane_off", C.c_int), ("mask_slope", C.c_float),
        ("tail_first_plane", C.c_int),
        ("out16", C.c_void_p), ("out16_planes_total", C.c_int), ("out16_plane_off", C.c_int),
        ("out16_up2", C.c_int), ("out16_pixel_shuffle", C.c_int),
        ("out32", C.c_void_p), ("out32_planes_total", C.c_int), ("out32_plane_off", C.c_int),
        ("out_nchw", C.c_void_p), ("out_nchw_c", C.c_int),
        ("tile_p", C.c_int), ("tile_mt", C.c_int),
        ("wpacked_rows", C.c_void_p), ("rows_nbn", C.c_int), ("rows_mode", C.c_int),
    ]


class WgradArgs(C.Structure):
    """mirror of esr_conv3x3_wgrad_args"""
    _fields_ = [
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("dtype", C.c_int),
        ("x", C.c_void_p), ("x_planes_total", C.c_int), ("x_plane_off", C.c_int),
        ("gy", C.c_void_p), ("gy_planes_total", C.c_int), ("gy_plane_off", C.c_int),
        ("cout", C.c_int), ("cin", C.c_int), ("lead", C.c_int),
        ("dw", C.c_void_p), ("db", C.c_void_p), ("scale", C.c_float), ("accumulate", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("cin_total", C.c_int), ("cin_off", C.c_int),
    ]


class PackItem(C.Structure):
    """mirror of esr_pack_item"""
    _fields_ = [("w", C.c_void_p), ("cout", C.c_int), ("cin", C.c_int), ("lead", C.c_int), ("kcp", C.c_int), ("dtype", C.c_int),
                ("transpose_flip", C.c_int), ("wpacked", C.c_void_p), ("bias_out", C.c_void_p), ("bias_in", C.c_void_p),
                ("wpacked_rows", C.c_void_p), ("rows_nbn", C.c_int)]


class AdamTensor(C.Structure):
    """mirror of esr_adam_tensor"""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_ulonglong)]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "esr_last_error": (C.c_char_p, []),
    "esr_version": (C.c_int, []),
    "esr_launch_count": (C.c_longlong, []),
    "esr_device_check": (C.c_int, []),
    "esr_set_deterministic": (C.c_int, [C.c_int]),
    "esr_debug_watchdog": (C.c_int, [C.POINTER(C.c_uint), C.c_int]),
    "esr_conv3x3_fwd": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "esr_conv3x3_fwd_batch": (C.c_int, [C.POINTER(ConvArgs), C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "esr_conv3x3_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "esr_conv3x3_packed_bytes_ex": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "esr_conv3x3_cin_planes": (C.c_int, [C.c_int, C.c_int]),
    "esr_pack_conv3x3_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_pack_batch_scratch_bytes": (C.c_size_t, [C.c_int]),
    "esr_pack_conv3x3_weights_batch": (C.c_int, [C.POINTER(PackItem), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]),
    "esr_conv3x3_rows_config": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "esr_conv3x3_rows_config_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "esr_pack_conv3x3_weights_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "esr_conv3x3_wgrad_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "esr_conv3x3_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "esr_sum_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "esr_pack_nchw": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "esr_pack_nchw_affine": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "esr_maxpool2x2_planes16": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "esr_maxpool2x2_bwd_planes16": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "esr_unpack_planes16": (C.c_int, [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]),
    "esr_unpack_planes32": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "esr_upsample2x_planes16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "esr_pixel_unshuffle2_planes16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "esr_latent_downscale": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "esr_downsum2x_planes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "esr_planes_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_planes_add_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_sep_adjoint_1d": (C.c_int, [C.c_void_p] + [C.c_int] * 12 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "esr_latent_grad": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "esr_cem_down": (C.c_int, [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_cem_inv": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p]),
    "esr_cem_up_add": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]),
    "esr_bn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "esr_bn_stats": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 7 +
                     [C.c_size_t, C.c_void_p]),
    "esr_bn_lrelu_fwd": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_int,
                                                                  C.c_void_p, C.c_void_p]),
    "esr_space_to_depth_planes16": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "esr_bn_lrelu_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 4 + [C.c_float, C.c_int, C.c_int, C.c_float,
                                   C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "esr_bn_dbl_workspace_bytes": (C.c_size_t, [C.c_int]),
    "esr_bn_tangent_fwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 4 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "esr_bn_double_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 6 +
                          [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_size_t, C.c_void_p]),
    "esr_linear_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_float, C.c_void_p, C.c_void_p]),
    "esr_linear_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_float, C.c_int] +
                       [C.c_void_p] * 4),
    "esr_structure_tensor_workspace_bytes": (C.c_size_t, [C.c_int]),
    "esr_structure_tensor_fwd": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "esr_structure_tensor_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "esr_soft_hist_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "esr_soft_hist_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_adam_scratch_bytes": (C.c_size_t, [C.c_int]),
    "esr_adam_multi": (C.c_int, [C.POINTER(AdamTensor), C.c_int, C.c_void_p, C.c_size_t] + [C.c_float] * 5 + [C.c_int, C.c_float, C.c_void_p]),
    "esr_l1_workspace_bytes": (C.c_size_t, []),
    "esr_l1_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "esr_l1_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esr_bce_rel_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "esr_bce_rel_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
}

NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3"]


def build(force=False, verbose=False):
    """Compile csrc/esr_api.cu into libesr_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(REPO_ROOT, "include", "esr_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "esr_api.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library (building it first if a compiler is present and it is stale/missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            build()
        else:
            raise EsrError("libesr_b200.so is missing and nvcc is not available: run __graft_entry__.build() first. "
                           "There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EsrError("esr_b200 error %d: %s" % (rc, load().esr_last_error().decode()))


def watchdog(reset=True):
    """(synchronising) returns the 8-word watchdog record; word 0 != 0 means a pipeline wait timed out."""
    buf = (C.c_uint * 8)()
    check(load().esr_debug_watchdog(buf, int(reset)))
    return list(buf)


def launch_count():
    return int(load().esr_launch_count())
