"""Flow side of the INDM hot path (drop-in for the reference's `flow_models` package on that path)."""
