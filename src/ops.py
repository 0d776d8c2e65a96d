"""src/ops.py of the reference holds the training losses (out of scope) and a duplicate of the IEF wrappers
(ops.py:184,270).  BASELINE.json names `src/ops.batch_orth_proj_idrot`; export it here as an alias."""
from src.tf_smpl.projection import batch_orth_proj_idrot  # noqa: F401
from src.models import call_hmr_ief, hmr_ief              # noqa: F401
