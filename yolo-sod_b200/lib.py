"""ctypes binding of the C-ABI shared library (include/ysod.h). The library is the product: if it is missing, or
there is no CUDA device, every entry point raises -- there is no CPU or PyTorch fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YSOD_LIB_PATH") or os.path.join(_HERE, "libysod.so")   # (override: A/B builds of the same sources, tools/gpu_abso.sh)
LIB_PATH_F16 = os.path.join(_HERE, "libysod_f16.so")   # the same sources built with -DYSOD_HALF=1 (IEEE fp16 storage, `half=True`)

F32, BF16 = 0, 1
CONV_UP2 = 0x40   # ysod.h YSOD_CONV_UP2
CONV_IMG_WEIGHTS = 0x80   # ysod.h YSOD_CONV_IMG_WEIGHTS
CONV_NO_SPLIT_STAGING = 0x10   # ysod.h YSOD_CONV_NO_SPLIT_STAGING
CONV_NO_PAIR = 0x04       # ysod.h YSOD_CONV_NO_PAIR
CONV_NO_STORE = 0x20      # ysod.h YSOD_CONV_NO_STORE
CONV_NO_DUO = 0x10000     # ysod.h YSOD_CONV_NO_DUO
STEM_INDIRECT = 0x10      # ysod.h YSOD_STEM_INDIRECT
ACT = {"none": 0, "silu": 1, "gelu": 2, "relu": 3, "sigmoid": 4, "hsigmoid": 5}

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); mirrors include/ysod.h one to one
PROTOTYPES = {
    "ysod_version": (i32, []),
    "ysod_last_error": (C.c_char_p, []),
    "ysod_compiled_arch": (i32, []),
    "ysod_storage_dtype": (i32, []),
    "ysod_nms_workspace_bytes": (i64, [i32, i32, i32, i32, i32]),
    "ysod_nms_batched": (i32, [vp, i32, i32, i32, f32, f32, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, i64, vp]),
    "ysod_nms_boxes_workspace_bytes": (i64, [i32]),
    "ysod_nms_boxes": (i32, [vp, vp, i32, f32, i32, vp, vp, vp, i64, vp]),
    "ysod_dfl_decode": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, f32, vp, i32, i32, vp]),
    "ysod_conv_tc_create": (i32, [C.POINTER(vp), vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32]),
    "ysod_conv_tc_create_ex": (i32, [C.POINTER(vp), vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32, i32]),
    "ysod_conv_tc_set_decode": (i32, [vp, vp, i32, i32, i32, f32]),
    "ysod_conv_tc_set_b2b": (i32, [vp, vp, vp, i32, i32, vp, i32, i32, i32, f32, vp, i32, i32]),
    "ysod_conv_tc_set_b2b_conv": (i32, [vp, vp, vp, i32]),
    "ysod_conv_tc_set_b2b_cat": (i32, [vp, vp, i32, vp, vp, i32, vp, i32]),
    "ysod_scale_weights": (i32, [vp, i32, i32, i32, vp, i32, vp, vp]),
    "ysod_conv_tc_run": (i32, [vp, vp]),
    "ysod_conv_tc_info": (i32, [vp, C.POINTER(i32)]),
    "ysod_conv_tc_destroy": (None, [vp]),
    "ysod_debug_trace": (i32, [vp, i32]),
    "ysod_conv_direct": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp]),
    "ysod_dwconv": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, i32, vp, i32, i32, vp]),
    "ysod_stem_conv": (i32, [vp, i32, i32, i32, vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    "ysod_stem_mma": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp]),
    "ysod_stem_mma_gap": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp]),
    "ysod_set_ptr": (i32, [vp, vp, vp]),
    "ysod_letterbox_u8": (i32, [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "ysod_scale_boxes": (i32, [vp, i32, i32, i32, vp, i32, vp]),
    "ysod_gap_partial": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "ysod_gap_gate": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp]),
    "ysod_se_gate": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
    "ysod_cbam_gate": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "ysod_scale_channels": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp]),
    "ysod_cbam_stats": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "ysod_cbam_apply": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, i32, vp]),
    "ysod_cbam_spatial": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, vp]),
    "ysod_ca_pool_workspace_floats": (i64, [i32, i32, i32, i32]),
    "ysod_ca_pool": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "ysod_ca_gate": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "ysod_ca_apply": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]),
    "ysod_sppf_pool": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp]),
    "ysod_upsample_copy": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "ysod_avgpool2d": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "ysod_glu": (i32, [vp, i32, i64, i32, i32, vp, i32, vp]),
    "ysod_upsample_add": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp, i32, vp]),
    "ysod_layernorm": (i32, [vp, i32, i64, i32, i32, vp, vp, f32, vp, i32, vp]),
    "ysod_window_partition_ln": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, f32, vp, vp, i32, vp]),
    "ysod_window_reverse": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "ysod_adaptive_pool_rows": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "ysod_bilinear_rows": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "ysod_dwconv3_ln": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp, f32, vp, i32, vp, i32, vp]),
    "ysod_mha_window_nhwc": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, f32, vp, i32, vp]),
    "ysod_swin64_fused": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]),
    "ysod_swin64_tc": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]),
    "ysod_swin64_tc_trace": (i32, [i32, vp]),
    "ysod_mha_core": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i64, i64, i64, f32, vp, i32, i64, vp]),
    "ysod_mha_core_ex": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i64, i64, i64, f32, vp, i32, i64, i32, vp]),
}

_libs = {}


class YsodError(RuntimeError):
    pass


def load(half: bool = False):
    """dlopen libysod.so (bf16 build; `half=True`: libysod_f16.so, the fp16 build of the same sources) and attach prototypes.
    Raises if the library has not been built (`__graft_entry__.build()`)."""
    key = bool(half)
    if key not in _libs:
        path = LIB_PATH_F16 if key else LIB_PATH
        if not os.path.exists(path):
            raise YsodError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                            "There is no CPU fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.ysod_storage_dtype() != (2 if key else 1):
            raise YsodError(f"{path}: built for storage dtype {lib.ysod_storage_dtype()}, expected {2 if key else 1}")
        _libs[key] = lib
    return _libs[key]


def check(rc, what="", half=False):
    if rc != 0:
        msg = load(half).ysod_last_error().decode("utf-8", "replace")
        raise YsodError(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise YsodError("no CUDA device: yolo-sod_b200 runs on sm_100a GPUs only (no CPU fallback)")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, offset_elems=0):
    """device pointer of a torch tensor (+ element offset)."""
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def call(name, *args, half=False):
    check(getattr(load(half), name)(*args), name, half)
