"""CPU oracle of the reference hot path -- test infrastructure only (see unet_oracle.py)."""
