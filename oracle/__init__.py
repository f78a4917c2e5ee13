"""CPU oracle for the KV-cache decode hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch fp32 CPU restatement of the reference's
algorithm for the hot path:

  * ``decoder_oracle``  - EnhancedTransformerDecoder memory builder, one-step
    KV-cache forward, generate_with_kv_cache, sample_for_reinforce
    (reference: src/superconductor/models/autoregressive_decoder.py:779-899,
    1175-1641).
  * ``encoder_oracle``  - FullMaterialsVAE encode / decode heads
    (reference: src/superconductor/models/attention_vae.py:625-822,
    src/superconductor/encoders/element_attention.py:73-214).
  * ``weights``         - seeded synthetic state_dicts with the reference's key
    names and shapes (no checkpoint ships with the reference).
  * ``vocab``           - FractionAwareTokenizer id layout and [5, V] type masks
    (reference: src/superconductor/tokenizer/fraction_tokenizer.py:130-338).
  * ``latent``          - slerp (reference: scripts/holdout/holdout_search.py:128-146).

Pinning: the reference holds no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, run in the build container by ``tests/golden/make_golden.py``
on the same seeded weights; the vectors are committed under ``tests/golden/``
and re-checked by ``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  The product package ``superconductor_vae_b200`` never
imports it and has no CPU fallback.
"""
