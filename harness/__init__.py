"""Synthetic model / train-step harness around the operator: not part of the product package."""
