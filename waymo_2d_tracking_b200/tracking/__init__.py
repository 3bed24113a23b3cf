"""Drop-in mirror of the reference's ``tracking/`` directory (SORT tracker CLI and library).

The reference's files are scripts that import each other by bare name (``from utils import ...``
with ``tracking/`` as the working directory, ``tracking/track.py:6``); this package accepts both
spellings: ``python -m waymo_2d_tracking_b200.tracking.track`` and, from inside this directory or
with it on ``PYTHONPATH``, ``python track.py`` / ``import utils``."""
