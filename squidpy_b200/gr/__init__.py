"""``squidpy_b200.gr`` — the spatial-statistics hot path of ``squidpy.gr`` on a B200."""

from ._build import (GridBuilder, KNNBuilder, RadiusBuilder, SpatialNeighborsResult, knn_2d, radius_2d, spatial_neighbors_grid,
                     spatial_neighbors_knn, spatial_neighbors_radius)
from ._ligrec import ligrec, ligrec_analysis
from ._nhood import NhoodEnrichmentResult, NhoodPlan, interaction_matrix, nhood_enrichment
from ._ppatterns import AutocorrPlan, co_occurrence, cooc_counts, spatial_autocorr
from ._ripley import pair_counts, ripley
from ._sepal import sepal, sepal_scores

__all__ = [
    "spatial_neighbors_knn",
    "spatial_neighbors_grid",
    "spatial_neighbors_radius",
    "SpatialNeighborsResult",
    "KNNBuilder",
    "GridBuilder",
    "RadiusBuilder",
    "knn_2d",
    "radius_2d",
    "ligrec",
    "sepal",
    "sepal_scores",
    "ligrec_analysis",
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
