"""fp32 CPU restatement of the EnhancedTransformerDecoder KV-cache decode path.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by the product.

Everything here is a *functional* restatement over a plain ``state_dict``
(no nn.Module), following these reference lines
(src/superconductor/models/autoregressive_decoder.py):

  build_memory           :779-873   (_create_memory / precompute_memory :875-899)
  step_hidden            :1196-1319 (_forward_one_step_with_cache) with
                         nn.TransformerDecoderLayer(norm_first=True, gelu) and
                         nn.MultiheadAttention semantics from torch.nn.functional
  decode_logits          :1413      (output_proj), :1417 (token_type_head),
                         :1439 (stop_head)
  generate_with_kv_cache :1321-1557 (mask :1416-1422, stop boost :1438-1457,
                         degenerate guard :1464-1466, entropy :1470-1482,
                         temperature :1485, top-k :1489, top-p :1494-1503,
                         argmax / multinomial :1506-1518, early exit :1547)
  sample_for_reinforce   :1559-1641

Pinning: checked against the reference modules themselves (imported from
/root/reference in the build container) by tests/golden/make_golden.py; the
resulting vectors live in tests/golden/ and tests/test_oracle_golden.py
re-checks them everywhere.  The reference has no tests of its own for this
path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

PAD_IDX, START_IDX, END_IDX = 0, 1, 2          # autoregressive_decoder.py:95-97 / tokenizer :47-49
HEADS_ORDER = ("tc_pred", "sc_pred", "hp_pred", "tc_class_logits", "competence",
               "element_count_pred", "family_composed_14")


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x):
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[prefix + ".bias"], 1e-5)


def infer_shape(sd: Dict[str, torch.Tensor]) -> dict:
    """Shape inference from key names / tensor shapes, the way the reference's loaders do
    (scripts/holdout/holdout_search.py:223-252)."""
    d = sd["token_embedding.weight"].shape[1]
    layers = 0
    while f"transformer_decoder.layers.{layers}.self_attn.in_proj_weight" in sd:
        layers += 1
    bottleneck = 0
    if "latent_to_memory.3.weight" in sd:                       # 0 Linear, 1 LN, 2 GELU, 3 Linear
        bottleneck = sd["latent_to_memory.0.weight"].shape[0]
        n_lat = sd["latent_to_memory.3.weight"].shape[0] // d
    else:
        n_lat = sd["latent_to_memory.2.weight"].shape[0] // d
    return dict(
        d_model=d, num_layers=layers, vocab_size=sd["token_embedding.weight"].shape[0],
        dim_feedforward=sd["transformer_decoder.layers.0.linear1.weight"].shape[0],
        max_len=sd["pos_encoding.pe"].shape[1], latent_dim=sd["latent_to_memory.0.weight"].shape[1],
        n_memory_tokens=n_lat, memory_bottleneck_dim=bottleneck,
        stoich_input_dim=sd["stoich_to_memory.0.weight"].shape[1] if "stoich_to_memory.0.weight" in sd else 0,
        n_stoich_tokens=(sd["stoich_to_memory.3.weight"].shape[0] // d) if "stoich_to_memory.3.weight" in sd else 0,
        heads_input_dim=sd["heads_to_memory.0.weight"].shape[1] if "heads_to_memory.0.weight" in sd else 0,
        heads_n_tokens=(sd["heads_to_memory.5.weight"].shape[0] // d) if "heads_to_memory.5.weight" in sd else 0,
        use_skip_connection="skip_to_memory.0.weight" in sd,
        skip_n_tokens=(sd["skip_to_memory.2.weight"].shape[0] // d) if "skip_to_memory.2.weight" in sd else 0,
    )


def heads_input_matrix(heads_pred: Dict[str, torch.Tensor], batch: int, family_dim: int = 14) -> torch.Tensor:
    """cat order tc, sc, hp, tc_class(5), competence, count, family(14) (:845-858)."""
    for name in HEADS_ORDER[:-1]:
        if heads_pred[name].shape[0] != batch:
            raise RuntimeError(f"heads_pred['{name}'] batch {heads_pred[name].shape[0]} != z batch {batch}")
    fam = heads_pred.get("family_composed_14")
    if fam is not None and fam.shape[0] != batch:
        raise RuntimeError(f"heads_pred['family_composed_14'] batch {fam.shape[0]} != z batch {batch}")
    parts = [heads_pred["tc_pred"].unsqueeze(-1), heads_pred["sc_pred"].unsqueeze(-1),
             heads_pred["hp_pred"].unsqueeze(-1), heads_pred["tc_class_logits"],
             heads_pred["competence"].unsqueeze(-1), heads_pred["element_count_pred"].unsqueeze(-1),
             fam if fam is not None else torch.zeros(batch, family_dim)]
    return torch.cat(parts, dim=-1)


def build_memory(sd, z, encoder_skip=None, stoich_pred=None, heads_pred=None, nhead: int = 8) -> torch.Tensor:
    """memory [B, M, d]: latent tokens, [skip], stoich, heads - in that order (:799-871)."""
    s = infer_shape(sd)
    B, d = z.shape[0], s["d_model"]
    if s["memory_bottleneck_dim"] > 0:
        h = F.gelu(_ln(sd, "latent_to_memory.1", _lin(sd, "latent_to_memory.0", z)))
        lat = _lin(sd, "latent_to_memory.3", h)
    else:
        lat = _lin(sd, "latent_to_memory.2", F.gelu(_lin(sd, "latent_to_memory.0", z)))
    parts = [lat.view(B, s["n_memory_tokens"], d)]
    if s["use_skip_connection"] and encoder_skip is not None:
        sk = _lin(sd, "skip_to_memory.2", F.gelu(_lin(sd, "skip_to_memory.0", encoder_skip)))
        parts.append(sk.view(B, s["skip_n_tokens"], d))
    if s["n_stoich_tokens"] > 0 and stoich_pred is not None:
        h = F.gelu(_ln(sd, "stoich_to_memory.1", _lin(sd, "stoich_to_memory.0", stoich_pred)))
        parts.append(_lin(sd, "stoich_to_memory.3", h).view(B, s["n_stoich_tokens"], d))
    if heads_pred is not None and s["heads_n_tokens"] > 0:
        hin = heads_input_matrix(heads_pred, B)
        h = F.gelu(_ln(sd, "heads_to_memory.1", _lin(sd, "heads_to_memory.0", hin)))
        h = F.gelu(_lin(sd, "heads_to_memory.3", h))
        parts.append(_lin(sd, "heads_to_memory.5", h).view(B, s["heads_n_tokens"], d))
    return torch.cat(parts, dim=1)


def _mha_heads(t, B, n, nhead):
    return t.view(B, n, nhead, -1).transpose(1, 2)             # [B, h, n, hd]


def step_hidden(sd, nhead, tok_emb, memory, cache: List[Dict[str, torch.Tensor]], position: int,
                trace: Optional[dict] = None, kv_round=None):
    """One decode step for one new token per row (:1196-1319).  ``cache`` is updated in place.

    ``kv_round``: optional callable applied to k, v before they are appended (the "bf16-KV oracle"
    variant of SURVEY.md section 7); None = the reference's fp32 cache.
    """
    pe = sd["pos_encoding.pe"]
    if position >= pe.shape[1]:
        raise IndexError(f"Position {position} exceeds PE buffer size {pe.shape[1]}.")
    B, d = tok_emb.shape[0], tok_emb.shape[-1]
    hd = d // nhead
    x = tok_emb + pe[:, position:position + 1, :]
    n_layers = len(cache)
    for li in range(n_layers):
        p = f"transformer_decoder.layers.{li}."
        # --- self attention over the cache
        xn = _ln(sd, p + "norm1", x)
        qkv = F.linear(xn, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        if kv_round is not None:
            k, v = kv_round(k), kv_round(v)
        cache[li]["key"] = torch.cat([cache[li]["key"], k], dim=1)
        cache[li]["value"] = torch.cat([cache[li]["value"], v], dim=1)
        n = cache[li]["key"].shape[1]
        qh = _mha_heads(q, B, 1, nhead)
        kh = _mha_heads(cache[li]["key"], B, n, nhead)
        vh = _mha_heads(cache[li]["value"], B, n, nhead)
        w = F.softmax(torch.matmul(qh, kh.transpose(-2, -1)) * (hd ** -0.5), dim=-1)
        a = torch.matmul(w, vh).transpose(1, 2).contiguous().view(B, 1, d)
        x = x + _lin(sd, p + "self_attn.out_proj", a)
        if trace is not None and li == 0:
            trace.setdefault("after_self_l0", []).append(x.clone())
        # --- cross attention to memory (nn.MultiheadAttention, q != k == v)
        xn = _ln(sd, p + "norm2", x)
        W, bvec = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]
        M = memory.shape[1]
        qc = _mha_heads(F.linear(xn, W[:d], bvec[:d]), B, 1, nhead)
        kc = _mha_heads(F.linear(memory, W[d:2 * d], bvec[d:2 * d]), B, M, nhead)
        vc = _mha_heads(F.linear(memory, W[2 * d:], bvec[2 * d:]), B, M, nhead)
        wc = F.softmax(torch.matmul(qc, kc.transpose(-2, -1)) * (hd ** -0.5), dim=-1)
        c = torch.matmul(wc, vc).transpose(1, 2).contiguous().view(B, 1, d)
        x = x + _lin(sd, p + "multihead_attn.out_proj", c)
        if trace is not None and li == 0:
            trace.setdefault("after_cross_l0", []).append(x.clone())
        # --- feed forward
        xn = _ln(sd, p + "norm3", x)
        x = x + _lin(sd, p + "linear2", F.gelu(_lin(sd, p + "linear1", xn)))
        if trace is not None and li == 0:
            trace.setdefault("after_l0", []).append(x.clone())
    return x                                                      # transformer_decoder.norm is None (:692-695)


def decode_logits(sd, x):
    """output_proj (LN, Linear, GELU, Linear) on the final hidden state (:698-704)."""
    return _lin(sd, "output_proj.4", F.gelu(_lin(sd, "output_proj.1", _ln(sd, "output_proj.0", x))))


def type_logits(sd, x):
    h = F.gelu(_lin(sd, "token_type_head.1", _ln(sd, "token_type_head.0", x)))
    return _lin(sd, "token_type_head.7", F.gelu(_lin(sd, "token_type_head.4", h)))


def stop_logit(sd, x):
    return _lin(sd, "stop_head.2", F.gelu(_lin(sd, "stop_head.0", x)))


def forward_teacher_forced(sd, nhead, z, target_tokens, encoder_skip=None, stoich_pred=None, cached_memory=None,
                           heads_pred=None):
    """EnhancedTransformerDecoder.forward with teacher_forcing_ratio = 1.0 (:901-985): all positions in parallel,
    causal mask plus key padding mask (input token == PAD, :952), no final norm.  Returns
    (logits [B,L,V], generated [B,L], stop_logits [B,L], type_logits [B,L,5], site_dup_logits [B,L] or None)."""
    memory = cached_memory if cached_memory is not None else build_memory(sd, z, encoder_skip, stoich_pred, heads_pred, nhead)
    inp = target_tokens[:, :-1]
    B, L = inp.shape
    pe = sd["pos_encoding.pe"]
    if L > pe.shape[1]:
        raise RuntimeError(f"sequence of {L} positions exceeds the PE buffer ({pe.shape[1]})")
    x = F.embedding(inp, sd["token_embedding.weight"]) + pe[:, :L, :]
    d = x.shape[-1]
    hd = d // nhead
    n_layers = infer_shape(sd)["num_layers"]
    pos = torch.arange(L)
    blocked = (pos[None, :] > pos[:, None])[None, None] | (inp == 0)[:, None, None, :]      # [B,1,L(query),L(key)]
    M = memory.shape[1]
    for li in range(n_layers):
        p = f"transformer_decoder.layers.{li}."
        xn = _ln(sd, p + "norm1", x)
        q, k, v = F.linear(xn, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]).chunk(3, dim=-1)
        qh, kh, vh = _mha_heads(q, B, L, nhead), _mha_heads(k, B, L, nhead), _mha_heads(v, B, L, nhead)
        sc = (torch.matmul(qh, kh.transpose(-2, -1)) * (hd ** -0.5)).masked_fill(blocked, float("-inf"))
        a = torch.matmul(F.softmax(sc, dim=-1), vh).transpose(1, 2).contiguous().view(B, L, d)
        x = x + _lin(sd, p + "self_attn.out_proj", a)
        xn = _ln(sd, p + "norm2", x)
        W, bvec = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]
        qc = _mha_heads(F.linear(xn, W[:d], bvec[:d]), B, L, nhead)
        kc = _mha_heads(F.linear(memory, W[d:2 * d], bvec[d:2 * d]), B, M, nhead)
        vc = _mha_heads(F.linear(memory, W[2 * d:], bvec[2 * d:]), B, M, nhead)
        wc = F.softmax(torch.matmul(qc, kc.transpose(-2, -1)) * (hd ** -0.5), dim=-1)
        x = x + _lin(sd, p + "multihead_attn.out_proj", torch.matmul(wc, vc).transpose(1, 2).contiguous().view(B, L, d))
        xn = _ln(sd, p + "norm3", x)
        x = x + _lin(sd, p + "linear2", F.gelu(_lin(sd, p + "linear1", xn)))
    logits = decode_logits(sd, x)
    dup = None
    if "site_dup_head.0.weight" in sd:
        dup = _lin(sd, "site_dup_head.2", F.gelu(_lin(sd, "site_dup_head.0", x))).squeeze(-1)
    return logits, logits.argmax(dim=-1), stop_logit(sd, x).squeeze(-1), type_logits(sd, x), dup


def generate_with_kv_cache(
    sd, nhead, z, encoder_skip=None, stoich_pred=None, temperature: float = 1.0,
    top_k: Optional[int] = None, top_p: Optional[float] = None, max_len: Optional[int] = None,
    return_log_probs: bool = False, return_entropy: bool = False, cached_memory=None,
    stop_boost: float = 0.0, hard_stop_threshold: float = 0.0, heads_pred=None, type_masks=None,
    site_dup_threshold: float = 0.0, *, generator: Optional[torch.Generator] = None,
    forced_tokens: Optional[torch.Tensor] = None, trace: Optional[dict] = None, kv_round=None,
    stop_when_all_finished: bool = True, max_steps: Optional[int] = None,
):
    """Same signature and semantics as the reference (:1321-1557).

    Extra keyword-only knobs for testing: ``generator`` for multinomial, ``forced_tokens`` [B, >=L]
    (teacher-forced replay: the forced token is emitted but log-prob / entropy are computed as
    usual), ``trace`` (dict that receives per-step logits etc.), ``kv_round``.
    """
    s = infer_shape(sd)
    seen = torch.zeros(z.shape[0] if z is not None else cached_memory.shape[0], s["vocab_size"], dtype=torch.bool) \
        if site_dup_threshold > 0 else None
    B = z.shape[0] if z is not None else cached_memory.shape[0]
    V = s["vocab_size"]
    max_len = max_len or s["max_len"]
    max_len = min(max_len, s["max_len"])
    memory = cached_memory if cached_memory is not None else build_memory(
        sd, z, encoder_skip, stoich_pred, heads_pred, nhead)
    cache = [{"key": torch.empty(B, 0, s["d_model"]), "value": torch.empty(B, 0, s["d_model"])}
             for _ in range(s["num_layers"])]
    cur = torch.full((B, 1), START_IDX, dtype=torch.long)
    finished = torch.zeros(B, dtype=torch.bool)
    toks, lps, ents = [], [], []
    for position in range(max_len - 1):
        emb = F.embedding(cur, sd["token_embedding.weight"])
        x = step_hidden(sd, nhead, emb, memory, cache, position, trace, kv_round)
        logits = decode_logits(sd, x).squeeze(1)
        if trace is not None:
            trace.setdefault("hidden", []).append(x.squeeze(1).clone())
            trace.setdefault("raw_logits", []).append(logits.clone())
        if type_masks is not None:
            tl = type_logits(sd, x).squeeze(1)
            pred_type = tl.argmax(dim=-1)
            logits = logits.masked_fill(~type_masks[pred_type], float("-inf"))
            if trace is not None:
                trace.setdefault("type_logits", []).append(tl.clone())
        if site_dup_threshold > 0 and position > 0:          # (:1424-1435) soft-suppress previously seen "elements"
            dl = _lin(sd, "site_dup_head.2", F.gelu(_lin(sd, "site_dup_head.0", x))).squeeze(1).squeeze(-1)
            suppress = torch.sigmoid(dl) < site_dup_threshold
            if trace is not None:
                trace.setdefault("site_dup_logit", []).append(dl.clone())
            if suppress.any():
                logits = logits.masked_fill(suppress.unsqueeze(1) & seen, -30.0)
        if stop_boost > 0:
            sl = stop_logit(sd, x).squeeze(1).squeeze(-1)
            sp = torch.sigmoid(sl)
            if trace is not None:
                trace.setdefault("stop_logit", []).append(sl.clone())
            logits[:, END_IDX] = logits[:, END_IDX] + stop_boost * sp
            if hard_stop_threshold > 0:
                force = (sp > hard_stop_threshold) & ~finished
                if force.any():
                    logits[force, :] = float("-inf")
                    logits[force, END_IDX] = 100.0
            if position > 10:
                logits[:, END_IDX] = logits[:, END_IDX] + 10.0 * (position - 10) / max(max_len - 10, 1)
        degenerate = bool(torch.isnan(logits).any() or torch.isinf(logits).any())
        if trace is not None:
            trace.setdefault("final_logits", []).append(logits.clone())
            trace.setdefault("degenerate", []).append(degenerate)
        if return_entropy:
            if degenerate:
                ents.append(torch.full((B,), math.log(max(V, 1))))
            else:
                pe_ = F.softmax(logits, dim=-1).clamp(min=1e-8)
                ents.append(-(pe_ * pe_.log()).sum(dim=-1))
        if temperature != 1.0:
            logits = logits / temperature
        if top_k is not None and top_k > 0:
            logits[logits < torch.topk(logits, top_k)[0][..., -1, None]] = float("-inf")
        if top_p is not None and top_p < 1.0:
            sl_, si_ = torch.sort(logits, descending=True)
            rm = torch.cumsum(F.softmax(sl_, dim=-1), dim=-1) > top_p
            rm[..., 1:] = rm[..., :-1].clone()
            rm[..., 0] = 0
            logits[rm.scatter(1, si_, rm)] = float("-inf")
        if temperature < 0.01:
            nxt = logits.argmax(dim=-1, keepdim=True)
            lp = torch.zeros(B)
            if forced_tokens is not None:
                nxt = forced_tokens[:, position:position + 1].to(torch.long)
        else:
            probs = F.softmax(logits, dim=-1)
            if degenerate:
                probs = torch.ones_like(probs) / probs.size(-1)
            if forced_tokens is not None:
                nxt = forced_tokens[:, position:position + 1].to(torch.long)
            else:
                nxt = torch.multinomial(probs, num_samples=1, generator=generator)
            lp = probs.clamp(min=1e-8).log().gather(1, nxt).squeeze(-1)
            if trace is not None:
                trace.setdefault("probs", []).append(probs.clone())
        toks.append(nxt)
        if return_log_probs:
            lps.append(lp)
        if seen is not None:
            # (:1525-1539) the id range 20..137 is the element range of the *pre-V13* vocabulary; kept as is
            tok_flat = nxt.squeeze(-1)
            is_elem = (tok_flat >= 20) & (tok_flat <= 137) & ~finished
            seen[torch.arange(seen.shape[0])[is_elem], tok_flat[is_elem]] = True
        finished = finished | (nxt.squeeze(-1) == END_IDX)
        cur = nxt
        if stop_when_all_finished and bool(finished.all()):
            break
        if forced_tokens is not None and position + 1 >= forced_tokens.shape[1]:
            break
        if max_steps is not None and position + 1 >= max_steps:      # bench.py: bounded CPU sample
            break
    generated = torch.cat(toks, dim=1)
    return (generated,
            torch.stack(lps, dim=1) if return_log_probs else None,
            torch.stack(ents, dim=1) if return_entropy else None)


def reinforce_mask(tokens: torch.Tensor) -> torch.Tensor:
    """1.0 up to and including the first END, all ones when a row has none (:1620-1639)."""
    B, L = tokens.shape
    is_end = tokens == END_IDX
    end_pos = torch.argmax(is_end.int(), dim=1)
    end_pos = torch.where(is_end.any(dim=1), end_pos, torch.tensor(L))
    return (torch.arange(L).unsqueeze(0).expand(B, -1) <= end_pos.unsqueeze(1)).float()


def sample_for_reinforce(sd, nhead, z, encoder_skip=None, stoich_pred=None, temperature: float = 0.8,
                         max_len=None, cached_memory=None, stop_boost: float = 0.0,
                         hard_stop_threshold: float = 0.0, heads_pred=None, type_masks=None,
                         site_dup_threshold: float = 0.0, **kw):
    tokens, lp, ent = generate_with_kv_cache(
        sd, nhead, z, encoder_skip=encoder_skip, stoich_pred=stoich_pred, temperature=temperature,
        max_len=max_len, return_log_probs=True, return_entropy=True, cached_memory=cached_memory,
        stop_boost=stop_boost, hard_stop_threshold=hard_stop_threshold, heads_pred=heads_pred,
        type_masks=type_masks, site_dup_threshold=site_dup_threshold, **kw)
    return tokens, lp, ent, reinforce_mask(tokens)


def first_end_lengths(tokens: torch.Tensor) -> torch.Tensor:
    """Number of positions a consumer reads per row: index of first END + 1, else L."""
    B, L = tokens.shape
    is_end = tokens == END_IDX
    pos = torch.argmax(is_end.int(), dim=1) + 1
    return torch.where(is_end.any(dim=1), pos, torch.tensor(L))


def scheduled_sampling_mask(batch: int, seq_len: int, ratio: float, positional: bool = False, decay: float = 0.5) -> torch.Tensor:
    """The keep-ground-truth mask of the reference's scheduled sampling (:1037-1045), drawn from the CURRENT global CPU
    generator exactly like the reference draws it (one torch.rand(B, L) call after the positions tensor is built)."""
    if positional and ratio < 1.0:
        pos = torch.arange(seq_len).float() / max(seq_len - 1, 1)
        tf_pos = (ratio * (1.0 + decay * (1.0 - pos))).clamp(0.0, 1.0)
        return torch.rand(batch, seq_len) < tf_pos.unsqueeze(0)
    return torch.rand(batch, seq_len) < ratio


def forward_scheduled_sampling(sd, nhead, z, target_tokens, use_gt_mask, encoder_skip=None, stoich_pred=None,
                               cached_memory=None, heads_pred=None):
    """EnhancedTransformerDecoder.forward with teacher_forcing_ratio < 1 (:987-1082): pass 1 on the ground truth, argmax,
    mix with the ground truth where `use_gt_mask` [B, L] is False, pass 2 on START + the mixed tokens."""
    memory = cached_memory if cached_memory is not None else build_memory(sd, z, encoder_skip, stoich_pred, heads_pred, nhead)
    first = forward_teacher_forced(sd, nhead, z, target_tokens, cached_memory=memory)
    mixed = torch.where(use_gt_mask, target_tokens[:, 1:], first[1])
    mixed_inputs = torch.cat([target_tokens[:, :1], mixed[:, :-1]], dim=1)
    # forward_teacher_forced drops the last column of what it is given: append a dummy one
    return forward_teacher_forced(sd, nhead, z, torch.cat([mixed_inputs, mixed_inputs[:, -1:]], dim=1), cached_memory=memory)
