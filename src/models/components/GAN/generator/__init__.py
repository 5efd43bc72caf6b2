"""Shim package: resolves the reference's Hydra _target_ strings to the B200 implementations (use_b200)."""
