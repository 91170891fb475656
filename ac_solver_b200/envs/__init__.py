from .ac_env import ACEnv, ACEnvConfig  # noqa: F401
from .ac_moves import ACMove, ac_moves_batch, concatenate_relators, conjugate  # noqa: F401
from .vector_env import ACVectorEnv  # noqa: F401
