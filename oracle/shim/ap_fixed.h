/* Empty stand-in so the reference's unit testbench (which includes the Xilinx headers
 * unconditionally, gnn_builder_lib_test/test.h:17-19) compiles in float mode with stock g++.
 * Nothing from Xilinx is reproduced here; fixed-point mode is not supported by this shim. */
#pragma once
#include <cassert>
#include <cmath>
