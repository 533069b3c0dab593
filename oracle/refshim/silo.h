/* silo.h — STAND-IN for the absent Silo dependency.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing is written to disk.  The arrays the reference's SiloWriter hands to DBPutQuadmesh /
 * DBPutQuadvar1 / DBPutQuadvar (src/SiloWriter.hpp:101-197: node coordinates, the owned quantity
 * and the cell-centred velocity it interpolates for output) are CAPTURED in memory so that the
 * test driver can compare them with the B200 output stage. */
#ifndef CFREF_SHIM_SILO_H
#define CFREF_SHIM_SILO_H

#include <string>
#include <vector>

struct DBfile
{
    std::string name;
};
struct DBoptlist
{
    int cycle = 0;
    double time = 0.0, dtime = 0.0;
};
typedef char const* const* DBCAS_t;

enum
{
    DB_CLOBBER = 0,
    DB_NOCLOBBER,
    DB_LOCAL,
    DB_PDB,
    DB_HDF5,
    DB_UNKNOWN,
    DB_APPEND,
    DB_READ,
    DB_ALL,
    DB_CARTESIAN,
    DB_ROWMAJOR,
    DB_COLMAJOR,
    DB_DOUBLE,
    DB_FLOAT,
    DB_COLLINEAR,
    DB_NONCOLLINEAR,
    DB_ZONECENT,
    DB_NODECENT,
    DB_QUADMESH,
    DB_QUADVAR,
    DBOPT_CYCLE,
    DBOPT_TIME,
    DBOPT_DTIME,
    DBOPT_COORDSYS,
    DBOPT_MAJORORDER
};

namespace cfref
{
struct SiloCapture
{
    int writes = 0; /* DBPutQuadmesh calls so far */
    int cycle = 0;
    double time = 0.0, dtime = 0.0;
    int ndims = 0;
    int node_dims[3] = { 0, 0, 0 };
    int zone_dims[3] = { 0, 0, 0 };
    std::vector<double> coords[3];
    std::vector<double> quantity;    /* "quantity": zone-centred, x fastest (LayoutLeft owned copy) */
    std::vector<double> velocity[3]; /* "velocity": cell-centred u, v */
    std::vector<std::string> multi_names;
};
inline SiloCapture& silo_capture()
{
    static SiloCapture c;
    return c;
}
} // namespace cfref

inline DBoptlist* DBMakeOptlist( int ) { return new DBoptlist(); }
inline int DBFreeOptlist( DBoptlist* o )
{
    delete o;
    return 0;
}
inline int DBAddOption( DBoptlist* o, int option, void* value )
{
    if ( option == DBOPT_CYCLE )
        o->cycle = *static_cast<int*>( value );
    else if ( option == DBOPT_TIME )
        o->time = *static_cast<double*>( value );
    else if ( option == DBOPT_DTIME )
        o->dtime = *static_cast<double*>( value );
    return 0;
}
inline DBfile* DBCreate( const char* name, int, int, const char*, int ) { return new DBfile{ name }; }
inline DBfile* DBOpen( const char* name, int, int ) { return new DBfile{ name }; }
inline int DBClose( DBfile* f )
{
    delete f;
    return 0;
}
inline int DBMkDir( DBfile*, const char* ) { return 0; }
inline int DBSetDir( DBfile*, const char* ) { return 0; }
inline void DBShowErrors( int, void ( * )( char* ) ) {}

/* coords: one array of node coordinates per dimension (DB_COLLINEAR) */
inline int DBPutQuadmesh( DBfile*, const char*, DBCAS_t, const void* coords_v, const int* dims, int ndims, int, int,
                          const DBoptlist* o )
{
    auto& c = cfref::silo_capture();
    double* const* coords = static_cast<double* const*>( coords_v );
    c.writes++;
    c.ndims = ndims;
    if ( o )
    {
        c.cycle = o->cycle;
        c.time = o->time;
        c.dtime = o->dtime;
    }
    for ( int d = 0; d < ndims; ++d )
    {
        c.node_dims[d] = dims[d];
        c.coords[d].assign( coords[d], coords[d] + dims[d] );
    }
    return 0;
}
inline int DBPutQuadvar1( DBfile*, const char*, const char*, const void* var, const int* dims, int ndims, const void*,
                          int, int, int, const DBoptlist* )
{
    auto& c = cfref::silo_capture();
    long n = 1;
    for ( int d = 0; d < ndims; ++d )
    {
        c.zone_dims[d] = dims[d];
        n *= dims[d];
    }
    const double* v = static_cast<const double*>( var );
    c.quantity.assign( v, v + n );
    return 0;
}
inline int DBPutQuadvar( DBfile*, const char*, const char*, int nvars, DBCAS_t, const void* vars_v, const int* dims,
                         int ndims, const void*, int, int, int, const DBoptlist* )
{
    auto& c = cfref::silo_capture();
    double* const* vars = static_cast<double* const*>( vars_v );
    long n = 1;
    for ( int d = 0; d < ndims; ++d )
        n *= dims[d];
    for ( int q = 0; q < nvars && q < 3; ++q )
        c.velocity[q].assign( vars[q], vars[q] + n );
    return 0;
}
inline int DBPutMultimesh( DBfile*, const char*, int n, char** names, int*, const DBoptlist* )
{
    auto& c = cfref::silo_capture();
    c.multi_names.assign( names, names + n );
    return 0;
}
inline int DBPutMultivar( DBfile*, const char*, int, char**, int*, const DBoptlist* ) { return 0; }

#endif
