"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/build_oracle.py).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds an input and the output of the reference's own sortByHost
(SourceCode/Baseline1.cu:15-64) for it, plus -- for 32 % nBits == 0 -- the output of the
reference's CPU restatement of its GPU algorithm (SourceCode/Baseline4.cu:67-273).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def main():
    assert O.ref_available("Baseline1"), "needs /root/reference"
    cases = {}
    # the reference's DEBUG test vector: n = 513, rand() & 0xFF, nBits = 4
    k = O.glibc_rand_keys(513, 0xFF)
    cases["kat_debug_n513_nbits4"] = (k, 4)
    # a prefix of the default test vector (rand(), nBits = 8) -- the full 2^24+1 case is
    # checked through its FNV fingerprint, not stored
    cases["kat_default_prefix_n4099_nbits8"] = (O.glibc_rand_keys(4099), 8)
    # full 32-bit range (the reference's rand() never sets bit 31), ragged sizes, odd widths
    cases["uniform_n1025_nbits8"] = (O.generate("uniform", 1025), 8)
    cases["uniform_n7681_nbits5"] = (O.generate("uniform", 7681, first=1000), 5)
    cases["uniform_n3000_nbits11"] = (O.generate("uniform", 3000, first=77), 11)
    cases["unique16_n5000_nbits4"] = (O.generate("unique16", 5000), 4)
    cases["zipf_n6000_nbits8"] = (O.generate("zipf", 6000), 8)
    cases["reversed_n2049_nbits2"] = (O.generate("reversed", 2049), 2)
    for name, (keys, nbits) in cases.items():
        out = O.ref_sort_by_host(keys, nbits)
        extra = {}
        if 32 % nbits == 0:
            extra["out_parallel_algorithm_b512"] = O.ref_sort_by_host_parallel_algorithm(keys, nbits, 512)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), keys=keys, nbits=np.int32(nbits),
                            out=out, **extra)
        print(name, keys.size, nbits, hex(O.fnv1a64(out)))


if __name__ == "__main__":
    main()
