// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for SUNDIALS' <nvector/nvector_serial.h>: the
// reference's src/cpp/janus_ode_common.hpp:2418-2434 defines a debug printer over an N_Vector, nothing on
// the RadauTe / evalDyns / evalJac path calls it.  Just enough declarations for that function to compile.
#pragma once
typedef long sunindextype;
typedef double sunrealtype;
struct _shim_N_Vector { sunindextype length; sunrealtype* data; };
typedef _shim_N_Vector* N_Vector;
inline sunindextype N_VGetLength(N_Vector v) { return v->length; }
inline sunrealtype* N_VGetArrayPointer(N_Vector v) { return v->data; }
