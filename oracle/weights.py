"""Seeded synthetic state_dicts and inputs (re-exported from the product package's ``synthetic`` module so
that bench.py's engine arm never has to import ``oracle``).  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from superconductor_vae_b200.synthetic import *  # noqa: F401,F403
from superconductor_vae_b200.synthetic import _bf16_round  # noqa: F401
