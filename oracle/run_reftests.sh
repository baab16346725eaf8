#!/bin/bash
# TEST INFRASTRUCTURE: run the reference's own 22 CTest programs (built
# unmodified by `make reftests`) from a scratch directory laid out like the
# reference's build tree (cwd = <scratch>/run, data in <scratch>/Resources).
# Needs /root/reference for the .dat fixtures; prints "name rc" per test.
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
SCR=$(mktemp -d /tmp/vr_reftests.XXXX)
mkdir -p "$SCR/run" "$SCR/Resources"
cp "$REF"/tests/Resources/*.dat "$SCR/Resources/"
cp "$REF"/tests/Resources/*.dat "$SCR/run/"
cp -n "$REF"/examples/*/*.dat "$SCR/run/" 2>/dev/null
fail=0
for exe in "$HERE"/_ref/t_*; do
  name=$(basename "$exe")
  (cd "$SCR/run" && timeout 300 "$exe" > "$SCR/$name.out" 2> "$SCR/$name.err")
  rc=$?
  echo "${name#t_} $rc"
  [ $rc -ne 0 ] && { fail=1; tail -3 "$SCR/$name.err"; }
done
rm -rf "$SCR"
exit $fail
