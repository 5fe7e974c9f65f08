"""Minimal `detectron2` surface imported by models/criterion.py:12-16."""
