"""Minimal stand-in for the absent third-party torch_robotics package (golden generation only)."""
