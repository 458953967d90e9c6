from lshrs_b200.storage.device import DeviceBucketStorage, DeviceIndex
from lshrs_b200.storage.memory import BucketOperation, BucketStorage, InMemoryStorage, bucket_key

__all__ = ["BucketOperation", "BucketStorage", "InMemoryStorage", "DeviceBucketStorage", "DeviceIndex", "bucket_key"]
