#!/bin/sh
# builds the host-side encoder model (test scaffolding) next to this script
cd "$(dirname "$0")" && g++ -O2 -std=c++17 -fPIC -shared -o libdeflate_model.so deflate_model.cc
