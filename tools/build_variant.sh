#!/bin/bash
# Build libvar_b200.so of another git revision for same-box A/B measurements:
#   tools/build_variant.sh <git-ref> <tag>   ->  var_b200/_variants/lib_<tag>.so   (select it with VAR_B200_LIB=...)
# Only valid while the C-ABI of the compared entry points is unchanged between the two revisions.
set -e
ref=$1; tag=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d /tmp/vb_variant_XXXX)
git -C "$root" archive "$ref" var_b200 include | tar -x -C "$tmp"
(cd "$tmp" && python -m var_b200.build > build.log 2>&1 || { tail -20 build.log; exit 1; })
mkdir -p "$root/var_b200/_variants"
cp "$tmp/var_b200/libvar_b200.so" "$root/var_b200/_variants/lib_$tag.so"
rm -rf "$tmp"
echo "built var_b200/_variants/lib_$tag.so from $ref"
