#!/bin/sh
# build_ref.sh — builds the UNMODIFIED reference (teoremma/pgen-rs, Rust) into oracle/_ref/pgen-rs so that the
# oracle restatement (oracle/pgen_oracle.c, oracle/oracle_np.py) and the CUDA path can be pinned against the real
# `pgen-rs filter` (oracle/pin_against_ref.py).  TEST INFRASTRUCTURE: nothing under pgen-rs_b200/ uses it.
#
# The reference is Rust with three un-vendored crates (clap 4.5.1, csv 1.3.0, evalexpr 11.3.0; Cargo.lock), so this
# needs `cargo` and either network access or a populated ~/.cargo registry.  Neither exists in the build image of
# this repository: there the script prints "parity unpinned" and exits 3, and every bit-exact claim in the docs is
# a claim against the restatement, not against the Rust binary.
#
#   exit 0  oracle/_ref/pgen-rs built
#   exit 3  cannot build here (no cargo / no reference checkout / crates not available): parity unpinned
#
# No reference SOURCE is copied into this repository: the build runs in a temporary copy of the checkout (the
# checkout is read-only) and only the binary lands in oracle/_ref/ (git-ignored).
set -u
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${PGB_REFERENCE:-/root/reference}
OUT="$HERE/_ref"
if ! command -v cargo >/dev/null 2>&1; then
    echo "parity unpinned: no cargo in PATH (the reference is Rust; nothing in $REF is C/C++)"
    exit 3
fi
if [ ! -f "$REF/Cargo.toml" ]; then
    echo "parity unpinned: no reference checkout at $REF (set PGB_REFERENCE)"
    exit 3
fi
TMP=$(mktemp -d "${TMPDIR:-/tmp}/pgb_ref_build.XXXXXX") || exit 3
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/." "$TMP/src" || exit 3
mkdir -p "$OUT"
# --locked: exactly the versions of Cargo.lock; try offline first (vendored/registry cache), then the network
if (cd "$TMP/src" && cargo build --release --locked --offline) >"$OUT/build.log" 2>&1 ||
   (cd "$TMP/src" && cargo build --release --locked) >>"$OUT/build.log" 2>&1; then
    cp "$TMP/src/target/release/pgen-rs" "$OUT/pgen-rs" || exit 3
    echo "built $OUT/pgen-rs"
    exit 0
fi
echo "parity unpinned: cargo could not build the reference (see $OUT/build.log): crates clap/csv/evalexpr unavailable offline?"
exit 3
