"""Optimisation utilities of the DiST fine-tuning step (reference: ``models/utils/{lr_policy,optimizer,losses}.py``)."""
from . import lr_policy, optimizer  # noqa: F401
