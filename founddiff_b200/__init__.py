"""founddiff_b200 — B200-native reverse-diffusion sampling path of FoundDiff."""
__version__ = "0.1.0"
