"""dada2_b200 -- B200-native (sm_100a CUDA) implementation of DADA2's dada() core loop behind the
reference's dada_uniques() interface.  See DESIGN.md / INTEGRATION.md."""
from .api import Dada2bError, Resident, dada_uniques  # noqa: F401
