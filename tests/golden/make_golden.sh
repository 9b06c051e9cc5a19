#!/bin/sh
# Regenerates tests/golden from the read-only reference checkout (run in the build container only).
set -e
REF=${1:-/root/reference/test/data}
HERE=$(dirname "$0")
cp "$REF/small.fasta.gz" "$REF/large.fasta.gz" "$HERE/"
for f in small.FNRout small.DIRout small.DIRout2 large.DIRout; do
  gzip -9 -n -c "$REF/$f.txt" > "$HERE/$f.txt.gz"
done
