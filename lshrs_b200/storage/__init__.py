from lshrs_b200.storage.memory import BucketOperation, BucketStorage, InMemoryStorage, bucket_key

__all__ = ["BucketOperation", "BucketStorage", "InMemoryStorage", "bucket_key"]
