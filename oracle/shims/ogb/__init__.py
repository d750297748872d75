"""ogb is only imported by data/eval code the oracle never loads."""
