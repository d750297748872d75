"""Imported, never used, by util/read_datasets.py:7."""
