# Stub a maintainer drops into IntEL/src/models/IntEL/ (INTEGRATION.md section 2): `--model_name IntEL_b200` selects the
# B200 path behind the reference's own model interface (main.py:127-130 resolves the class by name).
from intel_sigir2023_b200.IntEL import IntEL as _B200
from models.IntEL import IntEL as _ref


class IntEL_b200(_B200):
    # same reader / runner / flags / state_dict as models/IntEL/IntEL.py; batch construction stays the reference's
    Dataset = _ref.IntEL.Dataset
