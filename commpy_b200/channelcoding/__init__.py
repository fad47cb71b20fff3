"""Channel coding: same public names as commpy.channelcoding (commpy/channelcoding/__init__.py:65-71)."""
from .convcode import Trellis, conv_encode, viterbi_decode, viterbi_decode_batch, puncturing, depuncturing  # noqa: F401
from .interleavers import RandInterlv  # noqa: F401
from .ldpc import (get_ldpc_code_params, ldpc_bp_decode, ldpc_bp_decode_batch, triang_ldpc_systematic_encode,  # noqa: F401
                   write_ldpc_params, build_matrix)
from .turbo import turbo_encode, map_decode, turbo_decode, map_decode_batch, turbo_decode_batch  # noqa: F401
