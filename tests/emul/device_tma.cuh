// device_tma.cuh — HOST STAND-IN (tests/emul only) for the product's PTX wrappers (mbarrier / TMA).
//
// A TMA box load becomes a synchronous copy with the tensor map's out-of-bounds zero fill; it is complete
// when tma_load_3d returns, so the mbarrier calls have nothing left to do.  What this checks is the kernels'
// index arithmetic, tile / halo / ghost handling and the shared-memory layout they assume — NOT the
// asynchronous pipeline (stage reuse, barrier parities), which only a GPU run exercises.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

inline uint32_t smem_u32( const void* p ) { return cfb_emul::smem_addr( p ); }
inline void mbar_init( uint32_t, uint32_t ) {}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx( uint32_t, uint32_t ) {}
inline bool mbar_try_wait( uint32_t, uint32_t ) { return true; }
inline void mbar_wait( uint32_t, uint32_t ) {}
inline void prefetch_tmap( const CUtensorMap* ) {}
inline void tma_load_3d( uint32_t dst, const CUtensorMap* map, uint32_t, int x, int y, int z )
{
    double* out = static_cast<double*>( cfb_emul::smem_ptr( dst ) );
    const long long d0 = (long long)map->dim[0], d1 = (long long)map->dim[1], d2 = (long long)map->dim[2];
    const double* base = static_cast<const double*>( map->base );
    for ( unsigned bz = 0; bz < map->box[2]; ++bz )
        for ( unsigned by = 0; by < map->box[1]; ++by )
            for ( unsigned bx = 0; bx < map->box[0]; ++bx )
            {
                const long long X = x + (long long)bx, Y = y + (long long)by, Z = z + (long long)bz;
                double v = 0.0; // CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE: zeros
                if ( X >= 0 && X < d0 && Y >= 0 && Y < d1 && Z >= 0 && Z < d2 )
                    v = base[( Z * (long long)map->stride[1] + Y * (long long)map->stride[0] ) / 8 + X];
                out[( (size_t)bz * map->box[1] + by ) * map->box[0] + bx] = v;
            }
}
