"""dspeed_b200 -- B200-native (sm_100a) implementation of dspeed's ProcessingChain
block-execution hot path behind dspeed's own processor / chain / build_dsp interface.

See DESIGN.md.  Importing the package does not touch the GPU; the CUDA library is
loaded on first use and there is no CPU fallback."""

__version__ = "0.1.0"

from .errors import DSPError, DSPFatal, ProcessingChainError  # noqa: F401
