"""dada2_b200 -- B200-native (sm_100a CUDA) implementation of DADA2's dada() core loop behind the
reference's dada_uniques() interface.  See DESIGN.md / INTEGRATION.md."""
from . import api  # noqa: F401
from .api import Dada2bError, PackedCall, Resident, dada_uniques  # noqa: F401
from . import bimera  # noqa: F401,E402
from . import merge  # noqa: F401,E402
from . import derep  # noqa: F401,E402
from . import errmodel  # noqa: F401,E402
