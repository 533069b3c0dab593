/* pmpio.h — STAND-IN for Silo's PMPIO baton passing on one rank.  TEST INFRASTRUCTURE ONLY. */
#ifndef CFREF_SHIM_PMPIO_H
#define CFREF_SHIM_PMPIO_H

#include <mpi.h>

typedef enum
{
    PMPIO_READ = 0,
    PMPIO_WRITE = 1
} PMPIO_iomode_t;

typedef void* ( *PMPIO_CreateFileCallBack )( const char* fname, const char* nsname, void* udata );
typedef void* ( *PMPIO_OpenFileCallBack )( const char* fname, const char* nsname, PMPIO_iomode_t iomode,
                                           void* udata );
typedef void ( *PMPIO_CloseFileCallBack )( void* file, void* udata );

struct PMPIO_baton_t
{
    PMPIO_CreateFileCallBack create;
    PMPIO_OpenFileCallBack open;
    PMPIO_CloseFileCallBack close;
    void* udata;
};

static inline PMPIO_baton_t* PMPIO_Init( int, PMPIO_iomode_t, MPI_Comm, int, PMPIO_CreateFileCallBack c,
                                         PMPIO_OpenFileCallBack o, PMPIO_CloseFileCallBack cl, void* udata )
{
    return new PMPIO_baton_t{ c, o, cl, udata };
}
static inline int PMPIO_GroupRank( const PMPIO_baton_t*, int ) { return 0; }
static inline void* PMPIO_WaitForBaton( PMPIO_baton_t* b, const char* fname, const char* nsname )
{
    return b->create( fname, nsname, b->udata ); /* first (and only) rank of its group creates the file */
}
static inline void PMPIO_HandOffBaton( const PMPIO_baton_t* b, void* file ) { b->close( file, b->udata ); }
static inline void PMPIO_Finish( PMPIO_baton_t* b ) { delete b; }

#endif
