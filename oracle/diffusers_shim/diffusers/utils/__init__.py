"""ORACLE / TEST INFRASTRUCTURE ONLY.  ``diffusers.utils`` surface the reference touches:
``logging.get_logger`` (/root/reference/modules/pipeline.py:4,11) and ``BaseOutput``."""
import logging as _pylogging
from types import SimpleNamespace

from ..models.unet_2d_condition import BaseOutput  # noqa: F401

logging = SimpleNamespace(get_logger=_pylogging.getLogger)
