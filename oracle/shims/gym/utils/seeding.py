"""TEST INFRASTRUCTURE ONLY -- gym.utils.seeding.hash_seed / _int_list_from_bigint as used at
algos/multiagent/main.py:478-479 (gym 0.21 behaviour: sha512 of the seed string, first 8 bytes, big-endian)."""
import hashlib
import struct


def hash_seed(seed=None, max_bytes=8):
    h = hashlib.sha512(str(seed).encode("utf8")).digest()
    return _bigint_from_bytes(h[:max_bytes])


def _bigint_from_bytes(b):
    sizeof_int = 4
    padding = sizeof_int - len(b) % sizeof_int
    b += b"\0" * padding
    int_count = int(len(b) / sizeof_int)
    unpacked = struct.unpack("{}I".format(int_count), b)
    accum = 0
    for i, val in enumerate(unpacked):
        accum += 2 ** (sizeof_int * 8 * i) * val
    return accum


def _int_list_from_bigint(bigint):
    if bigint < 0:
        raise ValueError("Seed must be non-negative")
    elif bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints
