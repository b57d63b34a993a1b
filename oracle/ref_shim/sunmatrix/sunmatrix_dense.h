// TEST INFRASTRUCTURE (oracle/_ref build only).  Empty stand-in: the reference includes this un-vendored
// header (src/cpp/radaute.hpp:17-18, src/cpp/janus_ode_common.hpp:4-17) but nothing on the RadauTe /
// evalDyns / evalJac path uses a symbol from it.
#pragma once
