"""Channel coding: same public names as commpy.channelcoding (commpy/channelcoding/__init__.py:65-71)."""
from .convcode import Trellis, conv_encode, viterbi_decode, viterbi_decode_batch, puncturing, depuncturing  # noqa: F401
