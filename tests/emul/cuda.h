// cuda.h — HOST STAND-IN (tests/emul only): the driver-API types the product's headers mention.
#pragma once
struct CUtensorMap
{
    alignas( 64 ) char opaque[128];
};
