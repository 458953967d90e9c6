from lshrs_b200.io.parquet import iter_parquet_vectors

__all__ = ["iter_parquet_vectors"]
