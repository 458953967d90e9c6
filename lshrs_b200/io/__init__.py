from lshrs_b200.io.parquet import DEFAULT_PARQUET_BATCH_SIZE, iter_parquet_vectors
from lshrs_b200.io.postgres import DEFAULT_POSTGRES_BATCH_SIZE, iter_postgres_vectors

__all__ = ["iter_parquet_vectors", "iter_postgres_vectors", "DEFAULT_PARQUET_BATCH_SIZE",
           "DEFAULT_POSTGRES_BATCH_SIZE"]
