"""ctypes binding of libdtts.so (include/dtts.h).  No torch types cross the ABI: only raw device pointers,
sizes and the CUDA stream handle.  A missing library is a hard error -- there is no CPU / PyTorch fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtts.so")

ABI_VERSION = 5
DTTS_MAX_UPS = 8
DTTS_MAX_RB = 4
# dtts_status (include/dtts.h)
DTTS_OK, DTTS_ERR_BAD_ARG, DTTS_ERR_BAD_SHAPE, DTTS_ERR_MISSING_WEIGHT = 0, -1, -2, -3
DTTS_ERR_WORKSPACE_TOO_SMALL, DTTS_ERR_UNSUPPORTED_ARCH, DTTS_ERR_CUDA, DTTS_ERR_ALIGNMENT = -4, -5, -6, -7


class WeightEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("numel", C.c_uint64)]


class AcousticDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden", "n_heads", "enc_layers", "ffn_kernel", "ffn_filter", "dict_dim", "word_size", "pinyin_size",
        "dur_layers", "dur_kernel", "dur_chans", "frames_multiple", "latent", "dec_layers", "dec_kernel",
        "flow_hidden", "flow_kernel", "flow_blocks", "flow_layers", "n_mel", "language_zh", "precision", "s2pa_route",
        "model", "ph_size", "word_enc_layers", "rel_window")]


MODEL_DICT, MODEL_PORTASPEECH = 0, 1


class PsTextIn(C.Structure):
    _fields_ = [("txt_tokens", C.c_void_p), ("ph2word", C.c_void_p), ("B", C.c_int32), ("Tp", C.c_int32),
                ("Tw", C.c_int32)]


class PsTextOut(C.Structure):
    _fields_ = [("ph_encoder_out", C.c_void_p), ("word_encoder_out", C.c_void_p), ("dur", C.c_void_p),
                ("dur_int", C.c_void_p), ("ilens", C.c_void_p)]


class VocoderDesc(C.Structure):
    _fields_ = [("n_mel", C.c_int32), ("init_ch", C.c_int32), ("n_ups", C.c_int32), ("n_rb", C.c_int32),
                ("up_rates", C.c_int32 * DTTS_MAX_UPS), ("up_kernels", C.c_int32 * DTTS_MAX_UPS),
                ("rb_kernels", C.c_int32 * DTTS_MAX_RB), ("rb_dilations", (C.c_int32 * 3) * DTTS_MAX_RB),
                ("precision", C.c_int32)]


class TextIn(C.Structure):
    _fields_ = [("word_tokens", C.c_void_p), ("pron_modified", C.c_void_p), ("keys", C.c_void_p),
                ("values", C.c_void_p), ("key_map", C.c_void_p), ("pinyin", C.c_void_p), ("pinyin_map", C.c_void_p),
                ("B", C.c_int32), ("Tw", C.c_int32), ("Lk", C.c_int32), ("Lp", C.c_int32)]


class TextOut(C.Structure):
    _fields_ = [("word_encoder_out", C.c_void_p), ("dict_attn", C.c_void_p), ("pron_attn", C.c_void_p),
                ("dur", C.c_void_p), ("dur_int", C.c_void_p), ("ilens", C.c_void_p)]


class DictBankStruct(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("values", C.c_void_p), ("key_map", C.c_void_p), ("tok_offsets", C.c_void_p),
                ("pinyin", C.c_void_p), ("pinyin_map", C.c_void_p), ("pin_offsets", C.c_void_p),
                ("n_entries", C.c_int32)]


class TextInBank(C.Structure):
    _fields_ = [("word_tokens", C.c_void_p), ("pron_modified", C.c_void_p), ("dict_ids", C.c_void_p),
                ("B", C.c_int32), ("Tw", C.c_int32), ("Lk", C.c_int32), ("Lp", C.c_int32)]


# every symbol include/dtts.h declares: name -> (restype, argtypes)
_P, _I, _U64, _F = C.c_void_p, C.c_int32, C.c_uint64, C.c_float
SYMBOLS = {
    "dtts_abi_version": (C.c_int, []),
    "dtts_last_error": (C.c_char_p, []),
    "dtts_acoustic_create": (C.c_int, [C.POINTER(AcousticDesc), _P, _U64, C.POINTER(WeightEntry), _I, _P,
                                       C.POINTER(_P)]),
    "dtts_acoustic_destroy": (C.c_int, [_P]),
    "dtts_text_workspace_bytes": (_U64, [_P, _I, _I, _I, _I]),
    "dtts_text_encode": (C.c_int, [_P, C.POINTER(TextIn), C.POINTER(TextOut), _P, _U64, _P]),
    "dtts_text_bank_workspace_bytes": (_U64, [_P, _I, _I, _I, _I]),
    "dtts_text_encode_bank": (C.c_int, [_P, C.POINTER(DictBankStruct), C.POINTER(TextInBank), C.POINTER(TextOut), _P,
                                        _U64, _P]),
    "dtts_acoustic_status": (C.c_int, [_P, _P, _I]),
    "dtts_ps_text_workspace_bytes": (_U64, [_P, _I, _I, _I]),
    "dtts_ps_text_encode": (C.c_int, [_P, C.POINTER(PsTextIn), C.POINTER(PsTextOut), _P, _U64, _P]),
    "dtts_ps_attend_workspace_bytes": (_U64, [_P, _I, _I, _I, _I]),
    "dtts_ps_attend": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _U64, _P]),
    "dtts_length_regulate_scan": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, C.POINTER(C.c_int32), _P]),
    "dtts_length_regulate_fill": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "dtts_expand": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "dtts_decode_workspace_bytes": (_U64, [_P, _I, _I]),
    "dtts_decode_mel": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P, _U64, _P]),
    "dtts_vocoder_create": (C.c_int, [C.POINTER(VocoderDesc), _P, _U64, C.POINTER(WeightEntry), _I, _P,
                                      C.POINTER(_P)]),
    "dtts_vocoder_destroy": (C.c_int, [_P]),
    "dtts_vocode_workspace_bytes": (_U64, [_P, _I, _I]),
    "dtts_vocode": (C.c_int, [_P, _P, _I, _I, _P, _P, _U64, _P]),
    "dtts_vocode_lens": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _U64, _P]),
    "dtts_wav_to_pcm16": (C.c_int, [_P, _U64, _P, _P]),
    "dtts_pron_tokens": (C.c_int, [_P, _P, C.POINTER(DictBankStruct), _P, _I, _I, _I, _P, _P]),
    "dtts_vocoder_launch_count": (_U64, [_P]),
    "dtts_acoustic_launch_count": (_U64, [_P]),
    "dtts_debug_set_tc_fuse": (C.c_int, [_I]),
    "dtts_debug_set_acoustic_fuse": (C.c_int, [_I]),
    "dtts_debug_conv1d": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "dtts_debug_tc_conv1d": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P,
                                       _U64, _P]),
}

_lib = None


def load():
    """Loads libdtts.so (building is __graft_entry__.build()'s job).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `make -C dict_tts_b200/csrc` "
                           "(or __graft_entry__.build()); this engine has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dtts_abi_version() != ABI_VERSION:
        raise RuntimeError("libdtts ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().dtts_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libdtts {what} failed (status {rc}): {msg}")


def make_table(table):
    """[(name, offset, numel)] -> (ctypes array, keepalive list)."""
    arr = (WeightEntry * len(table))()
    keep = []
    for i, (name, off, n) in enumerate(table):
        b = name.encode("utf-8")
        keep.append(b)
        arr[i].name = b
        arr[i].offset = off
        arr[i].numel = n
    return arr, keep
