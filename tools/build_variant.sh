#!/bin/bash
# build variants/<name>.so from a git revision of the kernel sources (A/B runs on the GPU box: tools/gpu_ab.sh)
#   tools/build_variant.sh <name> [<git-rev>|WORK] [extra nvcc flags...]
set -e
name=$1; rev=${2:-HEAD}; shift; shift || true
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/bitnetmcu_b200" "$tmp/include" "$root/variants"
if [ "$rev" = WORK ]; then   # the working tree as it is
  cp -r "$root/bitnetmcu_b200/csrc" "$tmp/bitnetmcu_b200/csrc"; cp -r "$root/include/." "$tmp/include/"
  rm -f "$tmp"/bitnetmcu_b200/csrc/*.o
else
  git -C "$root" archive "$rev" bitnetmcu_b200/csrc include | tar -x -C "$tmp"
fi
make -C "$tmp/bitnetmcu_b200/csrc" EXTRA="$*" > "$tmp/build.log" 2>&1 || { tail -20 "$tmp/build.log"; exit 1; }
cp "$tmp/bitnetmcu_b200/libbitnetmcu_b200.so" "$root/variants/$name.so"
rm -rf "$tmp"
echo "built variants/$name.so from $rev"
