#!/bin/bash
# Packages the OSQP stand-in (osqp_shim/) as if it were an installed OSQP: <prefix>/include/osqp/osqp.h + <prefix>/lib/libosqp.so.
# Only there to exercise the OSQP_PREFIX switch of the Makefile in an image without OSQP:
#   bash fake_osqp_prefix.sh /tmp/fake_osqp && make OSQP_PREFIX=/tmp/fake_osqp OUT=/tmp/ref_with_prefix
set -e
p=${1:?usage: fake_osqp_prefix.sh PREFIX}
cd "$(dirname "$0")"
mkdir -p "$p/include/osqp" "$p/lib"
cp osqp_shim/*.h "$p/include/osqp/"
g++ -std=c++11 -O3 -DNDEBUG -fPIC -ffp-contract=off -w -Iosqp_shim -shared -o "$p/lib/libosqp.so" osqp_shim/osqp_shim.cpp
echo "$p"
