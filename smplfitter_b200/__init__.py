"""smplfitter_b200 -- Blackwell-native drop-in for the ``smplfitter.pt`` hot path.

``smplfitter_b200.pt`` exposes ``BodyModel`` / ``BodyFitter`` / ``BodyConverter`` with the
reference signatures (/root/reference/src/smplfitter/pt/__init__.py:25-32); all device
arithmetic runs in hand-written sm_100a CUDA behind the C ABI declared in
``include/smplfit_b200.h``.  There is no CPU or eager-PyTorch fallback.
"""

__version__ = '0.1.0'
