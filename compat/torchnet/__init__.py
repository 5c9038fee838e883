"""Minimal torchnet stand-in: the meters and the dataset wrapper learning/main.py and the dataset
adapters use (learning/main.py:183-185,221; learning/s3dis_dataset.py:57-62)."""
from . import dataset, meter  # noqa: F401
