#!/bin/bash
# builds the C++ API test against libtqgpu.so (plain g++: the host mirror needs no CUDA headers)
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O1 -I ../../include -I ../../triqs_b200/cpp test_gfs_api.cpp -o test_gfs_api -L ../../triqs_b200 -ltqgpu -Wl,-rpath,'$ORIGIN/../../triqs_b200'
echo "built tests/cpp/test_gfs_api"
