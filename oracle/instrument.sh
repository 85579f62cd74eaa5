#!/bin/sh
# oracle/instrument.sh REF OUTDIR -- TEST INFRASTRUCTURE.
# Generates OUTDIR/wurblpt/: every reference header symlinked, except bvh.hpp and wurblpt.hpp,
# which are regenerated from the reference with counter increments inserted (and nothing else
# changed). The generated files live under oracle/_ref/ (git-ignored) and are never committed.
set -e
REF="$1"
OUT="$2/wurblpt"
rm -rf "$OUT"
mkdir -p "$OUT"
for f in "$REF"/libwurblpt/*; do
    ln -s "$f" "$OUT/$(basename "$f")"
done
rm "$OUT/bvh.hpp" "$OUT/wurblpt.hpp"
sed -e 's|^#pragma once|#pragma once\n#include "ref_counters.hpp"|' \
    -e 's|^\( *\)HitRecord hr; // initialized with haveHit=false|\1WPT_REF_COUNT(calls);\n&|' \
    -e 's|^\( *\)const BVHNodeLinear& node = _nodes\[currentNodeIndex\];|&\n\1WPT_REF_COUNT(nodes);|' \
    -e 's|^\( *\)HitRecord currentHr = node.hitable->hit(|\1WPT_REF_COUNT(leaves);\n&|' \
    "$REF/libwurblpt/bvh.hpp" > "$OUT/bvh.hpp"
sed -e 's|^#pragma once|#pragma once\n#include "ref_counters.hpp"|' \
    -e 's#^\( *\)HitRecord \(hr\|directHR\|lightHR\) = scene.bvh().hit(#\1WPT_REF_COUNT(rays);\n&#' \
    "$REF/libwurblpt/wurblpt.hpp" > "$OUT/wurblpt.hpp"
grep -c WPT_REF_COUNT "$OUT/bvh.hpp" "$OUT/wurblpt.hpp"
