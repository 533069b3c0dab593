// Cajita_Partitioner.hpp — forwards to the single-file Cajita stand-in (oracle/refshim/Cajita.hpp).
// TEST INFRASTRUCTURE ONLY; see oracle/refshim/README.md.
#ifndef CFREF_SHIM_CAJITA_PARTITIONER_HPP
#define CFREF_SHIM_CAJITA_PARTITIONER_HPP
#include <Cajita.hpp>
#endif
