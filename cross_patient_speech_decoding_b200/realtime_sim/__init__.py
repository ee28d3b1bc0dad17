"""GPU-backed counterparts of the data-module hooks the reference's realtime simulation calls
on its hot path (realtime_sim/realtime_datamodule.py:813-894): PCA latent reduction and
source -> target alignment.  Tensors in, tensors out, as in the reference."""
from .hooks import align_to_target, reduce_to_latent_space  # noqa: F401
