"""Launched by tests/test_gpu_configs.py with torchrun on 2 GPUs (not collected by pytest): the product's multi-GPU
API (superconductor_vae_b200.parallel) over NCCL against the unsharded call.

  * generate_sharded on C512 (z, stoich_pred, heads_pred, type masks) and on a skip-connection model
    (encoder_skip passed by keyword: every per-row keyword must be cut to the rank's rows);
  * generate_sharded from cached_memory alone (z = None);
  * sample_for_reinforce_sharded: rows come back in the reference's global sample-major order."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import superconductor_vae_b200 as S                                   # noqa: E402
from superconductor_vae_b200 import parallel, synthetic as W          # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer  # noqa: E402

rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def first_end(t):
    is_end = t == 2
    return torch.where(is_end.any(dim=1), is_end.int().argmax(dim=1) + 1, torch.full((t.shape[0],), t.shape[1], device=t.device))


def same_up_to_end(a, b):
    L = min(a.shape[1], b.shape[1])
    n = first_end(b).clamp(max=L)
    live = torch.arange(L, device=a.device).unsqueeze(0) < n.unsqueeze(1)
    return bool(((a[:, :L] == b[:, :L]) | ~live).all())


ok = True
# ---- C512, greedy with masks + stop head
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)
N = 777                                                               # ragged split: 389 + 388 rows
z = W.make_latents(N, 2048, 5).to(dev)
st, hp = W.make_conditioning(N, 13, 5)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
kw = dict(temperature=0.001, max_len=64, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
full, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
sh, _, _ = parallel.generate_sharded(dec, z, stoich_pred=st, heads_pred=hp, **kw)
ok &= sh.dtype == torch.int16 and sh.shape[0] == N and same_up_to_end(sh.to(torch.int64), full)
mem = dec.precompute_memory(z, None, st, hp)
sh2, _, _ = parallel.generate_sharded(dec, None, cached_memory=mem, **kw)
ok &= same_up_to_end(sh2.to(torch.int64), full)

# ---- a model with the skip connection: encoder_skip is a per-row KEYWORD argument
shape = W.TINY_SKIP
dsk = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(shape, 0), nhead=shape.nhead, device=dev)
n2 = 37
z2 = W.make_latents(n2, shape.latent_dim, 9).to(dev)
st2, hp2 = W.make_conditioning(n2, shape.stoich_input_dim, 9)
st2, hp2 = st2.to(dev), {k: v.to(dev) for k, v in hp2.items()}
skip = torch.randn((n2, shape.encoder_skip_dim), generator=torch.Generator().manual_seed(3)).to(dev)
kw2 = dict(temperature=0.001, max_len=shape.max_len)
full2, _, _ = dsk.generate_with_kv_cache(z2, encoder_skip=skip, stoich_pred=st2, heads_pred=hp2, **kw2)
sh3, _, _ = parallel.generate_sharded(dsk, z2, stoich_pred=st2, heads_pred=hp2, encoder_skip=skip, **kw2)
ok &= same_up_to_end(sh3.to(torch.int64), full2)
no_skip, _, _ = dsk.generate_with_kv_cache(z2, stoich_pred=st2, heads_pred=hp2, **kw2)
ok &= not torch.equal(no_skip, full2)                                  # the skip tokens do change the decode

# ---- RLOO: base batch sharded, sample-major order restored
B, k = 64, 4
zb, stb, hpb = z[:B], st[:B], {n_: v[:B] for n_, v in hp.items()}
kw3 = dict(temperature=1.2, max_len=64, stop_boost=10.0)
t, lp, en, mk = parallel.sample_for_reinforce_sharded(dec, zb, k, stoich_pred=stb, heads_pred=hpb, _seed=3, **kw3)
t1, lp1, en1, mk1 = dec.sample_for_reinforce(zb.repeat(k, 1), stoich_pred=stb.repeat(k, 1),
                                             heads_pred={n_: v.repeat(k, *([1] * (v.dim() - 1))) for n_, v in hpb.items()},
                                             _seed=3, **kw3)
ok &= tuple(t.shape[:1]) == (B * k,) and t.dtype == torch.int64
# step-0 entropy depends on the latent only: the sharded rows must sit where the unsharded ones do
ok &= bool(torch.equal(en[:, 0], en1[:, 0]))
ok &= bool(torch.equal(mk, (torch.arange(t.shape[1], device=dev).unsqueeze(0) < first_end(t).unsqueeze(1)).float()))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if int(flag) == 1 else "SHARDED_MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
