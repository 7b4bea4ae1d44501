"""Input feed in front of the SA/FP path."""
