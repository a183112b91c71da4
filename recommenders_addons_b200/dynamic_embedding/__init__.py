"""Mirror of the reference's `tfra.dynamic_embedding` export list for the hot path
(/root/reference/tensorflow_recommenders_addons/dynamic_embedding/__init__.py:17-53)."""
from .table import (CuckooHashTable, CuckooHashTableConfig, CuckooHashTableCreator, HkvEvictStrategy, HkvHashTable,
                    HkvHashTableConfig, HkvHashTableCreator, KVCreator)
from .variable import (Variable, default_partition_fn, embedding_lookup, embedding_lookup_unique,
                       get_variable, unique)
from .ops import SparseIds, embedding_lookup_sparse, safe_embedding_lookup_sparse
from .optimizer import DynamicEmbeddingOptimizer, FusedAdagrad, FusedAdam
from .restrict_policies import FrequencyRestrictPolicy, RestrictPolicy, TimestampRestrictPolicy
from .sharded import PeerShardedVariable, ShardedVariable
from . import layers

__all__ = [
    "CuckooHashTable", "CuckooHashTableConfig", "CuckooHashTableCreator", "HkvEvictStrategy", "HkvHashTable", "HkvHashTableConfig",
    "HkvHashTableCreator", "KVCreator", "Variable", "default_partition_fn", "embedding_lookup",
    "embedding_lookup_unique", "get_variable", "unique", "SparseIds", "embedding_lookup_sparse",
    "safe_embedding_lookup_sparse", "DynamicEmbeddingOptimizer", "FusedAdagrad", "FusedAdam", "ShardedVariable", "PeerShardedVariable", "layers",
    "RestrictPolicy", "TimestampRestrictPolicy", "FrequencyRestrictPolicy",
]
