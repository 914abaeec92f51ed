class EmbeddingShardingPlanner:  # names only
    pass


class Topology:
    pass
