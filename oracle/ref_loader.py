"""Import the unmodified reference modules from oracle/_ref (built by oracle/make_ref.py) -- test infrastructure only.

`load()` returns the three reference classes on this path; it raises FileNotFoundError when oracle/_ref is absent (a
checkout that never ran the recipe), so callers can skip or fall back to the golden-pinned oracle port explicitly."""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "superconductor"))


def load():
    if not available():
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    sys.dont_write_bytecode = True
    with contextlib.redirect_stdout(io.StringIO()):            # the package prints an isotope-database banner on import
        from superconductor.models.attention_vae import FullMaterialsVAE
        from superconductor.models.autoregressive_decoder import EnhancedTransformerDecoder
        from superconductor.tokenizer.fraction_tokenizer import FractionAwareTokenizer
    return EnhancedTransformerDecoder, FullMaterialsVAE, FractionAwareTokenizer


def reference_decoder(shape, state_dict):
    """The reference's EnhancedTransformerDecoder built for `shape` (superconductor_vae_b200.synthetic.DecoderShape) with
    `state_dict` loaded strictly, in eval mode, on the CPU."""
    Dec, _, _ = load()
    dec = Dec(latent_dim=shape.latent_dim, d_model=shape.d_model, nhead=shape.nhead, num_layers=shape.num_layers,
              dim_feedforward=shape.dim_feedforward, max_len=shape.max_len, n_memory_tokens=shape.n_memory_tokens,
              encoder_skip_dim=shape.encoder_skip_dim, use_skip_connection=shape.use_skip_connection,
              vocab_size=shape.vocab_size, stoich_input_dim=shape.stoich_input_dim,
              memory_bottleneck_dim=shape.memory_bottleneck_dim)
    dec.load_state_dict(state_dict, strict=True)
    return dec.eval()
