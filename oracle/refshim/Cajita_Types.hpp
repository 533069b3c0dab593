// Cajita_Types.hpp — forwards to the single-file Cajita stand-in (oracle/refshim/Cajita.hpp).
// TEST INFRASTRUCTURE ONLY; see oracle/refshim/README.md.
#ifndef CFREF_SHIM_CAJITA_TYPES_HPP
#define CFREF_SHIM_CAJITA_TYPES_HPP
#include <Cajita.hpp>
#endif
