class Client:  # stub: the oracle never talks to BigQuery
    def __init__(self, *a, **k):
        raise RuntimeError("no BigQuery in the oracle harness")
