"""Stand-in for runx (NVIDIA experiment manager, requirements: runx==0.0.6) — logging to stdout only."""
