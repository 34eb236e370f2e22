from .sharded import RowWiseShardedEmbedding, ShardedDeepFM, allreduce_dense_grads, shard_rows  # noqa: F401
