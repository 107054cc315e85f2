"""ctypes binding of libs2i.so (C ABI: include/s2i.h).  There is NO CPU fallback: if the CUDA library
cannot be loaded every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libs2i.so")
_lib = None


class S2IError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    """Mirror of ``s2i_gemm_desc`` (include/s2i.h)."""
    _fields_ = [
        ("A", C.c_void_p), ("a_mn", C.c_int), ("aC", C.c_int), ("aW", C.c_int), ("aH", C.c_int), ("aB", C.c_int),
        ("a_sw", C.c_longlong), ("a_sh", C.c_longlong), ("a_sb", C.c_longlong), ("taps", C.c_int),
        ("a_c0", C.c_int), ("a_hoff", C.c_int), ("a_zmode", C.c_int),
        ("B", C.c_void_p), ("b_mn", C.c_int), ("bI", C.c_int), ("bR", C.c_int), ("bZ", C.c_int),
        ("b_sr", C.c_longlong), ("b_sz", C.c_longlong), ("b_c0", C.c_int), ("b_hoff", C.c_int), ("b_zmode", C.c_int),
        ("N", C.c_int), ("Kc", C.c_int), ("Z", C.c_int), ("zh", C.c_int), ("bf16", C.c_int), ("BN", C.c_int), ("splits", C.c_int),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int),
        ("residual", C.c_void_p), ("res_ld", C.c_longlong),
        ("out32", C.c_void_p), ("ld32", C.c_longlong), ("out16", C.c_void_p), ("ld16", C.c_longlong),
        ("out16_bf16", C.c_int), ("c_sb", C.c_longlong), ("c_sh", C.c_longlong), ("relu", C.c_int), ("qscale", C.c_float),
        ("out_glu", C.c_void_p), ("ld_glu", C.c_longlong), ("scratch32", C.c_void_p), ("b_static", C.c_int),
        ("colstat", C.c_void_p), ("colstat_ld", C.c_longlong), ("colstat_cap", C.c_int), ("colstat_bps", C.POINTER(C.c_int)),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.taps = 1
        self.aH = self.aB = self.bZ = self.Z = self.zh = 1
        self.alpha = 1.0
        for k, v in kw.items():
            setattr(self, k, v)


class UNetConfig(C.Structure):
    """Mirror of ``s2i_unet_config`` (include/s2i.h)."""
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("block_out_channels", C.c_int * 4),
                ("num_heads", C.c_int * 4), ("layers_per_block", C.c_int), ("cross_attention_dim", C.c_int),
                ("sample_size", C.c_int), ("ctx_len", C.c_int)]


class VaeConfig(C.Structure):
    """Mirror of ``s2i_vae_config`` (include/s2i.h)."""
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("latent_channels", C.c_int),
                ("block_out_channels", C.c_int * 4), ("layers_per_block", C.c_int)]


def lib():
    """Load (building in-tree first if only sources are present) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        from . import build as _build
        _build.build()
    try:
        h = C.CDLL(_LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise S2IError(f"cannot load {_LIB_PATH}: {e}. The sketch-guided path has no CPU fallback.") from e
    h.s2i_last_error.restype = C.c_char_p
    h.s2i_launch_count.restype = C.c_longlong
    h.s2i_profile_begin.argtypes = [C.c_void_p]
    h.s2i_profile_end.argtypes = [C.c_char_p, C.c_int]
    h.s2i_profile_set_peaks.argtypes = [C.c_double, C.c_double]
    h.s2i_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    h.s2i_gemm.restype = C.c_int
    h.s2i_gemm_set_tma_epilogue.argtypes = [C.c_int]
    h.s2i_gemm_force_msub.argtypes = [C.c_int]
    h.s2i_gemm_set_pair.argtypes = [C.c_int]
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p)
    h.s2i_attention.argtypes = [vp, C.c_longlong, C.c_int, vp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_longlong, vp, vp]
    h.s2i_attention_backward.argtypes = [vp, C.c_longlong, C.c_int, vp, C.c_longlong, C.c_int, C.c_int, vp, vp, C.c_longlong,
                                         vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp,
                                         C.c_longlong, C.c_int, vp, C.c_longlong, C.c_int, C.c_int, vp]
    h.s2i_unet_create.argtypes = [C.POINTER(UNetConfig), C.POINTER(vp)]
    h.s2i_unet_destroy.argtypes = [vp]
    h.s2i_unet_destroy.restype = None
    h.s2i_unet_load.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp), ip, C.POINTER(C.c_longlong)]
    h.s2i_unet_forward.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, C.c_int, vp]
    h.s2i_unet_tap.argtypes = [vp, C.c_int, fp, ip, ip, ip, ip]
    h.s2i_unet_tap_stride.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong)]
    L = C.c_longlong
    h.s2i_groupnorm_forward.argtypes = [vp, L, C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, C.c_int, vp, L, vp, L, vp, vp]
    h.s2i_groupnorm_forward_colstat.argtypes = [vp, L, C.c_int, C.c_int, C.c_int, vp, L, C.c_int, C.c_int, vp, vp, C.c_float, C.c_int,
                                                vp, L, vp, L, vp, vp]
    h.s2i_groupnorm_backward.argtypes = [vp, L, vp, L, C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, C.c_int, vp, vp, vp, L,
                                         vp, L, vp, L, vp]
    h.s2i_unet_backward.argtypes = [vp, C.POINTER(vp), vp, vp]
    h.s2i_unet_backward_samples.argtypes = [vp, C.POINTER(vp), vp, C.c_int, C.c_int, vp]
    h.s2i_unet_load_sat.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp), ip, C.POINTER(C.c_longlong)]
    h.s2i_unet_set_sat_feature.argtypes = [vp, C.c_char_p, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    h.s2i_unet_set_sat_scale.argtypes = [vp, C.c_float, vp]
    h.s2i_unet_debug.argtypes = [vp, C.c_int]
    h.s2i_unet_debug_get.argtypes = [vp, C.c_char_p, fp, C.POINTER(C.c_longlong), ip, ip, ip, ip]
    h.s2i_sketch_encoder_create.argtypes = [C.POINTER(UNetConfig), C.POINTER(vp)]
    h.s2i_sketch_encoder_forward.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp]
    h.s2i_sketch_encoder_num_res_samples.argtypes = [vp]
    h.s2i_sketch_encoder_res_sample.argtypes = [vp, C.c_int, fp, C.POINTER(C.c_longlong), ip, ip, ip, ip]
    h.s2i_vae_create.argtypes = [C.POINTER(VaeConfig), C.POINTER(vp)]
    h.s2i_vae_destroy.argtypes = [vp]
    h.s2i_vae_destroy.restype = None
    h.s2i_vae_load.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp), ip, C.POINTER(C.c_longlong)]
    h.s2i_vae_encode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    h.s2i_vae_decode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    h.s2i_vae_arena_bytes.argtypes = [vp]
    h.s2i_vae_arena_bytes.restype = C.c_longlong
    h.s2i_unet_arena_bytes.argtypes = [vp]
    h.s2i_unet_arena_bytes.restype = C.c_longlong
    ll, f = C.POINTER(C.c_longlong), C.c_float
    h.s2i_lgp_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    h.s2i_lgp_destroy.argtypes = [vp]
    h.s2i_lgp_destroy.restype = None
    h.s2i_lgp_load.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp), ip, ll]
    h.s2i_lgp_set_grad_rounding.argtypes = [vp, C.c_int]
    h.s2i_lgp_forward_taps.argtypes = [vp, C.POINTER(vp), ip, ip, C.c_int, C.c_int, vp, f, C.c_int, vp]
    h.s2i_lgp_forward_nchw.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    h.s2i_lgp_forward_taps_batch.argtypes = [vp, C.POINTER(vp), ip, ip, C.c_int, C.c_int, vp, vp]
    h.s2i_lgp_train_step.argtypes = [vp, vp, f, f, f, f, f, C.c_int, vp, vp]
    h.s2i_lgp_get_param.argtypes = [vp, C.c_char_p, vp, C.c_longlong]
    h.s2i_lgp_output.argtypes = [vp, vp, vp]
    h.s2i_lgp_loss_backward.argtypes = [vp, vp, C.POINTER(vp), vp, C.POINTER(f), vp]
    h.s2i_lgp_loss_backward_cond.argtypes = [vp, vp, C.POINTER(vp), vp, C.POINTER(f), vp]
    h.s2i_cfg_ddim_step.argtypes = [vp, vp, C.c_int, C.c_int, f, f, f, f, f, C.c_int, vp, vp]
    h.s2i_guidance_update.argtypes = [vp, vp, vp, C.c_int, C.c_int, f, vp, vp]
    h.s2i_sampler_create.argtypes = [vp, vp, C.POINTER(vp)]
    h.s2i_sampler_context_changed.argtypes = [vp]
    h.s2i_sampler_destroy.argtypes = [vp]
    h.s2i_sampler_destroy.restype = None
    h.s2i_sampler_step.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, f, f, f, f, f, f, C.c_int, C.c_int, f, f,
                                   C.c_int, vp, vp]
    h.s2i_cfg_dpmpp_step.argtypes = [vp, vp, vp, C.c_int, C.c_int, f, f, f, f, f, f, f, C.c_int, C.c_int, vp, vp]
    h.s2i_sampler_step_dpmpp.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, f, f, f, f, f, f, f, f, C.c_int, C.c_int,
                                         C.c_int, f, C.c_int, vp, vp]
    _lib = h
    return h


def check(rc):
    if rc != 0:
        raise S2IError(f"libs2i error {rc}: {lib().s2i_last_error().decode(errors='replace')}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(desc):
    check(lib().s2i_gemm(C.byref(desc), stream_ptr()))


def launch_count():
    return int(lib().s2i_launch_count())


def profile_begin(peak_tflops=0.0, peak_hbm_gbs=0.0):
    """Start per-launch device timing of every libs2i kernel on the current stream.  With the roofline denominators given,
    profile_end() also reports each class's time at its launches' own rooflines ("roof_ms")."""
    check(lib().s2i_profile_set_peaks(float(peak_tflops), float(peak_hbm_gbs)))
    check(lib().s2i_profile_begin(stream_ptr()))


def profile_end():
    """-> {kernel class: {"launches", "ms", "flops", "bytes", "roof_ms"}} since profile_begin()."""
    buf = C.create_string_buffer(1 << 16)
    n = lib().s2i_profile_end(buf, len(buf))
    if n < 0:
        check(n)
    out = {}
    for line in buf.value.decode().splitlines():
        tag, cnt, ms, fl, by, roof = line.split()
        out[tag] = {"launches": int(cnt), "ms": float(ms), "flops": float(fl), "bytes": float(by), "roof_ms": float(roof)}
    return out
