"""``squidpy_b200.gr`` — the spatial-statistics hot path of ``squidpy.gr`` on a B200."""

from ._nhood import NhoodEnrichmentResult, NhoodPlan, interaction_matrix, nhood_enrichment
from ._ppatterns import AutocorrPlan, co_occurrence, cooc_counts, spatial_autocorr
from ._ripley import pair_counts, ripley

__all__ = [
    "nhood_enrichment",
    "interaction_matrix",
    "spatial_autocorr",
    "co_occurrence",
    "ripley",
    "NhoodEnrichmentResult",
    "NhoodPlan",
    "AutocorrPlan",
    "cooc_counts",
    "pair_counts",
]
