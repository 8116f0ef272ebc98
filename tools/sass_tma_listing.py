"""profiles/r02_sass_tma_kernels.txt: the TMA / mbarrier instructions of every kernel of the shipped library.
Usage: cuobjdump -sass medpy_b200/lib/libmedpy_b200_gc.so | python tools/sass_tma_listing.py > profiles/r02_sass_tma_kernels.txt"""
import re
import sys

keep = re.compile(r"UTMALDG|UTMASTG|SYNCS|FENCE\.VIEW\.ASYNC|ELECT|UBLKCP")
counts, fn = {}, None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and keep.search(line):
        counts.setdefault(fn, []).append(re.sub(r"\s+/\*[0-9a-f]{16}\*/\s*$", "", line.rstrip()))
print("# SASS evidence for the TMA-staged kernels of libmedpy_b200_gc.so (sm_100a), round 2")
print("# cuobjdump -sass medpy_b200/lib/libmedpy_b200_gc.so | python tools/sass_tma_listing.py : lines with UTMALDG / SYNCS (mbarrier) /")
print("# FENCE.VIEW.ASYNC / ELECT per kernel; kernels without such instructions are omitted.  No UTMASTG: results are written with plain")
print("# coalesced stores (10 store streams of 256 B per warp and z-step in k_build_tile).")
print()
for k, v in counts.items():
    print("== %s  (UTMALDG x %d, SYNCS x %d)" % (k, sum("UTMALDG" in x for x in v), sum("SYNCS" in x for x in v)))
    for x in v[:24]:
        print(x)
    if len(v) > 24:
        print("        ... (%d more)" % (len(v) - 24))
    print()
