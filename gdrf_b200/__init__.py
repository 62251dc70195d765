"""gdrf_b200: B200-native ELBO + gradient of the sparse multinomial GDRF (san-soucie/gdrf hot path)."""
from . import _lib
from .checkpoint import intersect_dicts, load_checkpoint, make_checkpoint, strip_optimizer
from .elbo import (GDRFElbo, elbo_value_and_grads, elbo_value_and_grads_from_host, marginal_mean, marginal_moments,
                   perplexity_from_mean)
from .kernels import KERNEL_DICT, RBF, Exponential, Matern32, Matern52, RationalQuadratic
from .models import SparseMultinomialGDRF
from .streaming import (StreamingData, streaming_epoch, streaming_probabilities, streaming_selection,
                        streaming_window)
from .svi import SVI, ClippedAdam, FusedSVI, shard_bounds

__all__ = ["GDRFElbo", "elbo_value_and_grads", "elbo_value_and_grads_from_host", "marginal_mean", "marginal_moments",
           "perplexity_from_mean", "RBF", "Matern32",
           "Matern52", "Exponential", "RationalQuadratic", "KERNEL_DICT", "SparseMultinomialGDRF", "SVI", "FusedSVI", "ClippedAdam", "shard_bounds",
           "StreamingData", "streaming_epoch", "streaming_probabilities", "streaming_selection", "streaming_window", "make_checkpoint", "load_checkpoint",
           "strip_optimizer", "intersect_dicts", "_lib"]
