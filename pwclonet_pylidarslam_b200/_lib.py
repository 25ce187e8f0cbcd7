"""ctypes binding of libpwclo_b200.so (the C ABI declared in include/pwclo_b200.h).

There is deliberately NO fallback: if the shared library is missing or a CUDA tensor is not
supplied, the call raises.  A CPU or PyTorch-eager stand-in would void every parity claim.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpwclo_b200.so")
_lib = None

_vp, _i, _u, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_float


class LayerT(ctypes.Structure):
    """pwclo_layer_t of include/pwclo_b200.h"""
    _fields_ = [("w", ctypes.c_void_p), ("b", ctypes.c_void_p), ("cin", ctypes.c_int), ("cout", ctypes.c_int)]


_LP = ctypes.POINTER(LayerT)

# name -> argtypes  (must list every symbol declared in include/pwclo_b200.h; tests check this)
SIGNATURES = {
    "pwclo_furthest_point_sampling": [_vp, _i, _i, _i, _u, _vp, _vp],
    "pwclo_furthest_point_sampling_prefix": [_vp, _i, _i, _i, _u, _vp, _vp, _vp, _vp],
    "pwclo_gather_points": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pwclo_gather_points_grad": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pwclo_group_points": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "pwclo_group_points_grad": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "pwclo_ball_query": [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "pwclo_three_nn": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "pwclo_three_interpolate": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pwclo_three_interpolate_grad": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pwclo_knn": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pwclo_knn_sorted": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp],
    "pwclo_knn_presort": [_vp, _i, _i, _i, _vp, ctypes.c_size_t, _vp],
    "pwclo_knn_search": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pwclo_set_conv": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _LP, _i, _vp, _vp],
    "pwclo_pointwise_mlp": [ctypes.POINTER(_vp), ctypes.POINTER(_i), _i, _i, _LP, _i, _vp, _vp],
    "pwclo_cost_volume_1": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _LP, _LP, _LP, _vp, _vp],
    "pwclo_cost_volume_2": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _LP, _LP, _vp, _vp],
    "pwclo_pose_head": [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "pwclo_set_conv_tc": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _LP, _i, _vp, _vp],
    "pwclo_pointwise_mlp_tc": [ctypes.POINTER(_vp), ctypes.POINTER(_i), _i, _i, _LP, _i, _vp, _vp],
    "pwclo_cost_volume_1_tc": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _LP, _LP, _LP, _vp, _vp],
    "pwclo_cost_volume_2_tc": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _LP, _LP, _vp, _vp],
    "pwclo_tc_selftest": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "pwclo_gather_rows3": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "pwclo_transpose": [_vp, _i, _i, _i, _i, _vp, _vp],
    "pwclo_prepare_scans": [_vp, _vp, _i, ctypes.c_longlong, _vp, _i, _vp, ctypes.c_ulonglong, _i, _vp, _vp, _vp, _vp,
                            ctypes.c_size_t, _vp],
    "pwclo_pose_to_matrix": [_vp, _i, _i, _i, _vp, _vp],
    "pwclo_accumulate_poses": [_vp, _i, _vp, _vp, _vp],
    "pwclo_prepare_scans_crop": [_vp, _vp, _i, ctypes.c_longlong, _vp, _i, _vp, ctypes.c_ulonglong, _i, _i, _i, ctypes.c_double,
                                 _i, _i, ctypes.c_double, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp],
    "pwclo_adam_step": [_vp, _vp, _vp, _vp, ctypes.c_size_t, _i, _f, _f, _f, _f, _f, _f, _vp],
    "pwclo_adam_step_dev": [_vp, _vp, _vp, _vp, ctypes.c_size_t, _vp, _vp, _f, _f, _f, _f, _f, _vp],
    "pwclo_bn_relu_train_fwd": [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pwclo_bn_relu_train_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pwclo_warp_fwd": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "pwclo_warp_bwd": [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "pwclo_conv1x1_small": [_vp, _vp, _i, _i, _i, _i, ctypes.c_longlong, _vp, _vp],
    "pwclo_conv1x1_wgrad": [_vp, _vp, _i, _i, _i, ctypes.c_longlong, _vp, _vp, _vp],
    "pwclo_cost_geometry_fwd": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "pwclo_cost_geometry_bwd": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "pwclo_maxpool_lastdim_fwd": [_vp, ctypes.c_longlong, _i, _vp, _vp, _vp],
    "pwclo_maxpool_lastdim_bwd": [_vp, _vp, ctypes.c_longlong, _i, _vp, _vp],
    "pwclo_softmax_pool_fwd": [_vp, _vp, ctypes.c_longlong, _i, _vp, _vp],
    "pwclo_softmax_pool_bwd": [_vp, _vp, _vp, ctypes.c_longlong, _i, _vp, _vp, _vp],
    "pwclo_pose_loss": [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
}


class PwcloError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PwcloError(
                f"{LIB_PATH} is missing: build it with `python -m pwclonet_pylidarslam_b200.csrc.build` "
                "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.pwclo_knn_workspace_bytes.argtypes = [_i, _i, _i]
        L.pwclo_knn_workspace_bytes.restype = ctypes.c_size_t
        L.pwclo_conv1x1_wgrad_workspace_bytes.argtypes = [_i, _i, _i, ctypes.c_longlong]
        L.pwclo_conv1x1_wgrad_workspace_bytes.restype = ctypes.c_size_t
        L.pwclo_bn_relu_workspace_bytes.argtypes = [_i, _i, _i]
        L.pwclo_bn_relu_workspace_bytes.restype = ctypes.c_size_t
        L.pwclo_prepare_scans_workspace_bytes.argtypes = [ctypes.c_longlong, _i]
        L.pwclo_prepare_scans_workspace_bytes.restype = ctypes.c_size_t
        L.pwclo_version.restype = ctypes.c_char_p
        L.pwclo_error_string.restype = ctypes.c_char_p
        L.pwclo_error_string.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        raise PwcloError(f"{what} failed: {lib().pwclo_error_string(code).decode()} (code {code})")


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
