"""Modules named exactly like the reference's (``neural_dynamics``, ``torchdiffeq``) so that the
unmodified scripts import this backend; put THIS directory first on sys.path (ndcn_b200.run does)."""
