"""ctypes binding of libscvae_b200.so (the C ABI declared in include/scvae_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or the device is
not a B200-class (sm_100) GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscvae_b200.so")
CSRC = os.path.join(_HERE, "csrc")

_lib = None


class EngineError(RuntimeError):
    pass


class DecoderConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_model", "nhead", "num_layers", "dim_feedforward", "vocab_size", "pe_len", "latent_dim",
        "n_memory_tokens", "memory_bottleneck_dim", "stoich_input_dim", "n_stoich_tokens", "heads_input_dim",
        "heads_n_tokens", "encoder_skip_dim", "skip_n_tokens")]


class GenerateArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_memory", C.c_int32), ("max_len", C.c_int32), ("memory", C.c_void_p),
        ("temperature", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float),
        ("stop_boost", C.c_float), ("hard_stop_threshold", C.c_float), ("site_dup_threshold", C.c_float),
        ("type_masks", C.c_void_p), ("want_log_probs", C.c_int32), ("want_entropy", C.c_int32),
        ("flags", C.c_uint32), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("out_tokens", C.c_void_p), ("out_log_probs", C.c_void_p), ("out_entropy", C.c_void_p),
        ("out_steps", C.POINTER(C.c_int32)), ("forced_tokens", C.c_void_p), ("memory_rows", C.c_int32)]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("seq_len", C.c_int32), ("n_memory", C.c_int32), ("memory", C.c_void_p),
        ("tokens", C.c_void_p), ("ld_tokens", C.c_int32), ("out_logits", C.c_void_p), ("out_stop", C.c_void_p),
        ("out_type", C.c_void_p), ("out_dup", C.c_void_p), ("flags", C.c_uint32)]


class EncoderConfig(C.Structure):
    _fields_ = [
        ("n_element_rows", C.c_int32), ("element_embed_dim", C.c_int32), ("n_attention_heads", C.c_int32),
        ("max_elements", C.c_int32), ("magpie_dim", C.c_int32), ("fusion_dim", C.c_int32), ("latent_dim", C.c_int32),
        ("n_encoder_hidden", C.c_int32), ("encoder_hidden", C.c_int32 * 4),
        ("n_decoder_hidden", C.c_int32), ("decoder_hidden", C.c_int32 * 4)]


class RewardConfig(C.Structure):
    """scv_reward_config (include/scvae_b200.h): GPURewardConfig + GPURewardConfigV14 fields under their reference names."""
    _fields_ = (
        [(n, C.c_float) for n in ("exact_match", "near_exact_1", "near_exact_2", "near_exact_3", "token_correct",
                                  "token_penalty", "length_mismatch_penalty", "fraction_digit_penalty",
                                  "fraction_structure_penalty")]
        + [("use_semantic_digit_penalty", C.c_int32)]
        + [(n, C.c_float) for n in ("semantic_digit_scale", "length_only_base_reward", "length_only_per_extra",
                                    "length_only_floor")]
        + [("v14", C.c_int32), ("use_continuous_reward", C.c_int32)]
        + [(n, C.c_float) for n in ("max_reward", "sharpness", "element_error_penalty", "integer_error_penalty",
                                    "fraction_error_penalty", "special_error_penalty", "too_short_base_reward",
                                    "too_short_per_missing", "too_short_floor")]
        + [("use_phased_curriculum", C.c_int32), ("reward_phase", C.c_int32), ("phase3_sharpness", C.c_float)]
        + [(n, C.c_int32) for n in ("v14_element_start", "v14_element_end", "v14_integer_start", "v14_integer_end",
                                    "v14_fraction_start")])


class ConstraintConfig(C.Structure):
    """scv_constraint_config (include/scvae_b200.h): VocabConfig + ConstraintRewardConfig + FamilyConstraintConfig."""
    _fields_ = (
        [(n, C.c_int32) for n in ("element_start", "element_end", "digit_start", "digit_end", "lparen_idx", "rparen_idx",
                                  "slash_idx", "pad_idx", "end_idx", "use_semantic_fractions", "fraction_token_start",
                                  "a1_enabled", "a2_enabled", "a4_enabled", "a7_enabled", "family_enabled")]
        + [(n, C.c_double) for n in ("a1_penalty", "a2_penalty_per_violation", "a4_penalty", "a7_penalty",
                                     "confidence_threshold")]
        + [("b_penalty", C.c_double * 8)])


HEADS_OUT_FIELDS = (
    "tc_pred", "magpie_pred", "attended_input", "tc_class_logits", "competence", "fraction_pred",
    "element_count_pred", "hp_pred", "sc_pred", "family_coarse_logits", "family_cuprate_sub_logits",
    "family_iron_sub_logits", "family_composed_14", "stoich_pred", "heads_input")


class EncoderHeadsOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in HEADS_OUT_FIELDS]


FLAG_H2_UNIFORM_FALLBACK = 1
FLAG_SYNC_EVERY_STEP = 2
FLAG_COMPACT_FINISHED = 4
FORWARD_NO_KEY_PADDING = 1

# name -> (restype, argtypes); also the list the "exports every declared symbol" test walks
SIGNATURES = {
    "scv_abi_version": (C.c_int, []),
    "scv_last_error": (C.c_char_p, []),
    "scv_launch_count": (C.c_int64, []),
    "scv_tune": (C.c_int, [C.c_char_p, C.c_int32]),
    "scv_trace_begin": (C.c_int, [C.c_int32]),
    "scv_trace_read": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "scv_profile_begin": (C.c_int, []),
    "scv_profile_end": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double)]),
    "scv_profile_category_name": (C.c_char_p, [C.c_int32]),
    "scv_decoder_create": (C.c_int, [C.POINTER(DecoderConfig), C.POINTER(C.c_void_p)]),
    "scv_decoder_destroy": (None, [C.c_void_p]),
    "scv_decoder_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scv_decoder_missing_weights": (C.c_int, [C.c_void_p]),
    "scv_decoder_build_memory": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "scv_decoder_generate": (C.c_int, [C.c_void_p, C.POINTER(GenerateArgs), C.c_void_p]),
    "scv_decoder_debug_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "scv_encoder_create": (C.c_int, [C.POINTER(EncoderConfig), C.POINTER(C.c_void_p)]),
    "scv_encoder_destroy": (None, [C.c_void_p]),
    "scv_encoder_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scv_encoder_missing_weights": (C.c_int, [C.c_void_p]),
    "scv_encoder_encode": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scv_encoder_heads": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(EncoderHeadsOut), C.c_void_p]),
    "scv_slerp_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "scv_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scv_greedy_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "scv_tokens_canonical_hash": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "scv_element_similarity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "scv_reward_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "scv_constraint_rewards": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "scv_op_linear": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_void_p]),
    "scv_op_tiled_elems": (C.c_int64, [C.c_int32, C.c_int32]),
    "scv_op_pack_tiled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "scv_op_split_tile_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "scv_op_split_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "scv_op_linear_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "scv_op_pack_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "scv_op_layernorm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_void_p]),
}


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) or f == "Makefile")
    for f in files + [os.path.join("..", "..", "include", "scvae_b200.h")]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def build(verbose: bool = False, clean: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into libscvae_b200.so (nvcc cross-compiles without a GPU).

    `clean=True` removes every object first, so a successful return proves that every source compiled (about 30 s).
    Either way the build is recorded in csrc/build/BUILD_INFO.json together with the hash of the sources it was made
    from; a library whose recorded hash differs from the current sources is rebuilt from scratch."""
    import json
    import time
    info_path = os.path.join(CSRC, "build", "BUILD_INFO.json")
    cur = _source_hash()
    try:
        with open(info_path) as fh:
            stale = json.load(fh).get("source_sha256") != cur
    except (OSError, ValueError):
        stale = True
    if clean or stale or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", CSRC, "clean"], capture_output=True, text=True)
        clean = True
    t0 = time.time()
    r = subprocess.run(["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True)
    if r.returncode != 0:
        raise EngineError("building libscvae_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    compiled = [ln.split(" -c ")[1].split()[0] for ln in r.stdout.splitlines() if " -c " in ln]
    if clean or compiled:
        os.makedirs(os.path.dirname(info_path), exist_ok=True)
        with open(info_path, "w") as fh:
            json.dump({"source_sha256": cur, "clean_build": bool(clean), "compiled": compiled, "seconds": round(time.time() - t0, 1),
                       "flags": "-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"}, fh, indent=1)
    if verbose:
        print(f"libscvae_b200.so: {len(compiled)} translation units compiled ({'clean' if clean else 'incremental'} build)")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises EngineError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or `make -C {CSRC}`); this package has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().scv_last_error().decode("utf-8", "replace")
        raise EngineError(f"{what}: {msg}" if what else msg)


def tune(**kv) -> None:
    """Set run-time launch tunables (include/scvae_b200.h scv_tune), e.g. tune(attn_ctas_per_sm=3, subbatches=2)."""
    for k, v in kv.items():
        check(lib().scv_tune(k.encode(), int(v)), f"tune({k})")


def trace_begin(max_records: int = 1 << 20) -> None:
    check(lib().scv_trace_begin(int(max_records)), "trace_begin")


def trace_read(max_records: int = 1 << 20):
    """numpy structured array (t0, t1 in ns, sm, kid, bx, by) of the CTAs traced since trace_begin()."""
    import numpy as np
    dt = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("sm", "<u4"), ("kid", "<u4"), ("bx", "<u4"), ("by", "<u4")])
    buf = np.zeros(max_records, dtype=dt)
    n = C.c_int32(0)
    check(lib().scv_trace_read(buf.ctypes.data_as(C.c_void_p), int(max_records), C.byref(n)), "trace_read")
    return buf[:n.value]


def launch_count() -> int:
    return int(lib().scv_launch_count())


def profile_begin() -> None:
    check(lib().scv_profile_begin(), "profile_begin")


def profile_end() -> dict:
    """{category: {launches, ms, flops, bytes}} accumulated since profile_begin()."""
    n = 16
    counts, ms, fl, by = (C.c_int32 * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    check(lib().scv_profile_end(n, counts, ms, fl, by), "profile_end")
    out = {}
    for i in range(n):
        name = lib().scv_profile_category_name(i).decode()
        if name and counts[i] > 0:
            out[name] = {"launches": int(counts[i]), "ms": float(ms[i]), "flops": float(fl[i]), "bytes": float(by[i])}
    return out


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name: str):
    if t is not None and not t.is_cuda:
        raise EngineError(f"{name} must live on a CUDA device (sm_100a); this package has no CPU fallback")
