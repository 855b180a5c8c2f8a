"""`Koopman` package mirror (reference: Koopman/koopmanEDMDc.py) — scoring and simulation run on the B200 engine."""
