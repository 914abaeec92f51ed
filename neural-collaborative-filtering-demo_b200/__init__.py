"""b200-ncf: the AdvancedNCF training + scoring hot path on B200 (sm_100a).

Import as `ncf_b200` (the directory name carries a hyphen, so the repo root holds a thin
`ncf_b200/` alias whose __path__ points here).
"""
from ._lib import NcfError, load as load_library  # noqa: F401
from .architecture import AdvancedNCF, CategoryHierarchy, MultiHeadAttention, TemporalEncoding  # noqa: F401
from .kjt import KeyedJaggedTensor, make_kjt  # noqa: F401
from .data_prep import InteractionSampler, first_appearance_index, remap_cardnumber, remap_product_id  # noqa: F401
from .metrics import calculate_metrics  # noqa: F401
from .scoring import CatalogueScorer, get_recommendations  # noqa: F401
from .trainer import ModelTrainer, NCFTrainEngine  # noqa: F401
from .sharding import ShardedCatalogueScorer, ShardedNCFEngine, ShardRouter  # noqa: F401
from .export import CosineIndex, export_product_embeddings, product_embedding_records  # noqa: F401

__all__ = ["AdvancedNCF", "MultiHeadAttention", "TemporalEncoding", "CategoryHierarchy", "KeyedJaggedTensor",
           "make_kjt", "NcfError", "load_library", "calculate_metrics", "CatalogueScorer", "get_recommendations",
           "ModelTrainer", "NCFTrainEngine", "ShardedNCFEngine", "ShardRouter", "ShardedCatalogueScorer", "InteractionSampler", "first_appearance_index",
           "remap_cardnumber", "remap_product_id", "CosineIndex", "export_product_embeddings", "product_embedding_records"]
