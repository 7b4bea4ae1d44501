"""Compatibility shims that let the reference's own model.py run unmodified on this engine."""
