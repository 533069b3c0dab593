// output.cu — the output stage that follows the hot path in Solver::solve (src/Solver.hpp:156,170-173;
// SURVEY.md §8f rank 2): what the reference's SiloWriter hands to Silo for one block,
//   src/SiloWriter.hpp:109-123  node coordinates of the owned cells,
//   src/SiloWriter.hpp:136-156  the owned quantity (ghosts dropped),
//   src/SiloWriter.hpp:172-186  the cell-centred velocity (Interpolation::interpolateVelocity<D,1> at
//                               LocalMesh::coordinates( Cell ) of every owned cell),
// re-designed so that writing does not stall the time loop: ONE extraction kernel (the reference: two
// kernels + three blocking deep copies) packs q and the D velocity components into a dense device
// buffer on the compute stream; the device -> pinned-host copy runs on its own stream behind an
// event; the files are written by the host when the NEXT write (or a flush) comes, i.e. while the GPU
// has been running the steps in between.  Silo/PMPIO are absent: the container is one .npy per variable
// and block under <dir>/raw/ plus a .json master per step written by rank 0 (the role of
// writeMultiObjects, src/SiloWriter.hpp:292-346), with the reference's file-name pattern (:379-384).
#include "cfb_internal.h"
#include "device_geo.cuh"

#include <cerrno>
#include <cstring>
#include <sys/stat.h>

namespace
{

struct OutputArgs
{
    const double* cur[4]; // Current q, u, v, w
    double* q;            // [nz][ny][nx]
    double* vel;          // [D][nz][ny][nx]
};

template <int D>
__global__ void __launch_bounds__( 256 )
    output_extract_kernel( const __grid_constant__ Geo g, const __grid_constant__ OutputArgs a )
{
    const int nx = g.n[0], ny = g.n[1];
    const long long ncell = (long long)nx * ny * g.n[2];
    for ( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ncell;
          t += (long long)gridDim.x * blockDim.x )
    {
        const int i = (int)( t % nx );
        const int j = (int)( ( t / nx ) % ny );
        const int k = (int)( t / ( (long long)nx * ny ) );
        const int idx[3] = { i, j, k };
        a.q[t] = a.cur[0][geo_off( g, i, j, k )]; // :149-153
        double x[3], v[3];
        geo_coordinates<D>( g, 0, idx, x );       // :179
        interp_velocity<D>( g, a.cur, x, v );     // :180-181
#pragma unroll
        for ( int d = 0; d < D; ++d )
            a.vel[(long long)d * ncell + t] = v[d]; // :182-183
    }
}

struct Pending
{
    bool active = false;
    std::string dir;
    int step = 0;
    double time = 0, dt = 0;
};

} // namespace

struct OutputStage
{
    double* d_buf = nullptr; // (1 + D) * ncell doubles
    double* h_buf = nullptr; // pinned mirror
    size_t elems = 0;
    // its own stream: the side stream of the ctx carries the NCCL halo traffic of the CG iterations,
    // which must not queue behind a multi-GB copy
    cudaStream_t io_stream = nullptr;
    cudaEvent_t ev_extracted = nullptr, ev_copied = nullptr;
    bool in_flight = false; // a copy has been enqueued and not yet waited for
    Pending pending;        // a file write waiting for that copy
    std::string solve_dir;  // cfb_set_output_dir
};

namespace
{

size_t ncell_of( const Geo& g ) { return (size_t)g.n[0] * g.n[1] * g.n[2]; }

int stage_init( cfb_ctx* c )
{
    if ( c->out )
        return CFB_OK;
    OutputStage* o = new OutputStage();
    c->out = o;
    return CFB_OK;
}

int stage_buffers( cfb_ctx* c )
{
    OutputStage* o = c->out;
    if ( o->d_buf )
        return CFB_OK;
    o->elems = ncell_of( c->g ) * ( 1 + c->g.D );
    CFB_CUDA( c, cudaMalloc( &o->d_buf, o->elems * sizeof( double ) ) );
    CFB_CUDA( c, cudaMallocHost( &o->h_buf, o->elems * sizeof( double ) ) );
    CFB_CUDA( c, cudaStreamCreateWithFlags( &o->io_stream, cudaStreamNonBlocking ) );
    CFB_CUDA( c, cudaEventCreateWithFlags( &o->ev_extracted, cudaEventDisableTiming ) );
    CFB_CUDA( c, cudaEventCreateWithFlags( &o->ev_copied, cudaEventDisableTiming ) );
    return CFB_OK;
}

// extraction kernel on the compute stream, copy on the side stream; returns without waiting
int extract_async( cfb_ctx* c )
{
    OutputStage* o = c->out;
    int rc = stage_buffers( c );
    if ( rc )
        return rc;
    const Geo& g = c->g;
    // The reference does not gather before writing (its ghost samples carry weight 0 up to rounding);
    // with several blocks the ghosts are refreshed so that the output does not depend on the
    // decomposition.  Ghost layers are rewritten by the next step's gather anyway.
    if ( c->cfg.use_nccl )
    {
        rc = cfb_gather( c, CFB_CURRENT );
        if ( rc )
            return rc;
    }
    OutputArgs a{};
    for ( int e = 0; e <= g.D; ++e )
        a.cur[e] = field_ptr( c, e, CFB_CURRENT );
    const long long ncell = (long long)ncell_of( g );
    a.q = o->d_buf;
    a.vel = o->d_buf + ncell;
    long long blocks = ( ncell + 255 ) / 256;
    const long long cap = (long long)c->sm_count * 16;
    const int grid = (int)( blocks < 1 ? 1 : ( blocks > cap ? cap : blocks ) );
    if ( g.D == 2 )
        output_extract_kernel<2><<<grid, 256, 0, c->stream>>>( g, a );
    else
        output_extract_kernel<3><<<grid, 256, 0, c->stream>>>( g, a );
    c->stats.kernel_launches += 1;
    CFB_CUDA( c, cudaGetLastError() );
    CFB_CUDA( c, cudaEventRecord( o->ev_extracted, c->stream ) );
    CFB_CUDA( c, cudaStreamWaitEvent( o->io_stream, o->ev_extracted, 0 ) );
    CFB_CUDA( c, cudaMemcpyAsync( o->h_buf, o->d_buf, o->elems * sizeof( double ), cudaMemcpyDeviceToHost,
                                  o->io_stream ) );
    CFB_CUDA( c, cudaEventRecord( o->ev_copied, o->io_stream ) );
    o->in_flight = true;
    return CFB_OK;
}

int wait_copy( cfb_ctx* c )
{
    OutputStage* o = c->out;
    if ( !o->in_flight )
        return CFB_OK;
    CFB_CUDA( c, cudaEventSynchronize( o->ev_copied ) );
    o->in_flight = false;
    return CFB_OK;
}

// SiloWriter.hpp:109-123: coordinates( Node(), {0,..,i,..,0} )[d] for the owned cells' nodes
void node_coordinates( const Geo& g, int d, double* out )
{
    for ( int i = 0; i <= g.n[d]; ++i )
        out[i] = g.ghost_low[d] + (double)( i + g.h ) * g.celld[d];
}

bool make_dir( const std::string& p )
{
    if ( mkdir( p.c_str(), 0777 ) == 0 || errno == EEXIST )
        return true;
    return false;
}

// numpy .npy version 1.0, little-endian float64, C order
bool write_npy( const std::string& path, const double* data, const std::vector<long long>& shape )
{
    std::string hdr = "{'descr': '<f8', 'fortran_order': False, 'shape': (";
    size_t count = 1;
    for ( size_t i = 0; i < shape.size(); ++i )
    {
        hdr += std::to_string( shape[i] ) + ( shape.size() == 1 || i + 1 < shape.size() ? "," : "" );
        if ( i + 1 < shape.size() )
            hdr += " ";
        count *= (size_t)shape[i];
    }
    hdr += "), }";
    // magic (6) + version (2) + header length (2) + header, padded with spaces to a multiple of 64, '\n' last
    size_t total = 10 + hdr.size() + 1;
    size_t pad = ( 64 - total % 64 ) % 64;
    hdr += std::string( pad, ' ' );
    hdr += '\n';
    FILE* f = std::fopen( path.c_str(), "wb" );
    if ( !f )
        return false;
    const unsigned char magic[8] = { 0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0 };
    const unsigned short hl = (unsigned short)hdr.size();
    const unsigned char hlb[2] = { (unsigned char)( hl & 0xff ), (unsigned char)( hl >> 8 ) };
    bool ok = std::fwrite( magic, 1, 8, f ) == 8 && std::fwrite( hlb, 1, 2, f ) == 2 &&
              std::fwrite( hdr.data(), 1, hdr.size(), f ) == hdr.size() &&
              std::fwrite( data, sizeof( double ), count, f ) == count;
    ok = ( std::fclose( f ) == 0 ) && ok;
    return ok;
}

std::string block_name( int rank, int step, const char* var )
{
    char buf[128];
    // src/SiloWriter.hpp:382-384: "data/raw/CajitaFluidsOutput%05d%05d.%s" (group rank, time step)
    std::snprintf( buf, sizeof( buf ), "raw/CajitaFluidsOutput%05d%05d.%s.npy", rank, step, var );
    return buf;
}

void partition_of( int n, int nb, int b, int& owned, int& offset )
{
    int base = n / nb, rem = n % nb;
    owned = base + ( b < rem ? 1 : 0 );
    offset = b * base + ( b < rem ? b : rem );
}

// the files of one write: this block's variables + (rank 0) the master that names every block
int write_files( cfb_ctx* c )
{
    OutputStage* o = c->out;
    const Pending& p = o->pending;
    const Geo& g = c->g;
    const int D = g.D, rank = c->cfg.world_rank;
    if ( !make_dir( p.dir ) || !make_dir( p.dir + "/raw" ) )
        return cfb_fail( c, CFB_ERR_INVALID, "cannot create output directory " + p.dir + ": " + std::strerror( errno ) );
    const long long ncell = (long long)ncell_of( g );
    std::vector<long long> cs, vs;
    vs.push_back( D );
    for ( int d = D - 1; d >= 0; --d )
    {
        cs.push_back( g.n[d] );
        vs.push_back( g.n[d] );
    }
    bool ok = write_npy( p.dir + "/" + block_name( rank, p.step, "quantity" ), o->h_buf, cs ) &&
              write_npy( p.dir + "/" + block_name( rank, p.step, "velocity" ), o->h_buf + ncell, vs );
    const char* nn[3] = { "nodes_x", "nodes_y", "nodes_z" };
    for ( int d = 0; d < D && ok; ++d )
    {
        std::vector<double> nodes( g.n[d] + 1 );
        node_coordinates( g, d, nodes.data() );
        ok = write_npy( p.dir + "/" + block_name( rank, p.step, nn[d] ), nodes.data(), { (long long)g.n[d] + 1 } );
    }
    if ( ok && rank == 0 )
    {
        // writeMultiObjects (src/SiloWriter.hpp:292-346): one master per step naming the blocks
        char name[64];
        std::snprintf( name, sizeof( name ), "/CajitaFluids%05d.json", p.step ); // :379-380
        FILE* f = std::fopen( ( p.dir + name ).c_str(), "w" );
        ok = f != nullptr;
        if ( f )
        {
            const cfb_config& cfg = c->cfg;
            std::fprintf( f, "{\"cycle\": %d, \"time\": %.17g, \"dtime\": %.17g, \"dim\": %d,\n", p.step, p.time, p.dt, D );
            std::fprintf( f, " \"global_num_cell\": [" );
            for ( int d = 0; d < D; ++d )
                std::fprintf( f, "%d%s", cfg.global_num_cell[d], d + 1 < D ? ", " : "],\n" );
            std::fprintf( f, " \"layout\": \"C order, x fastest: quantity[(z,) y, x], velocity[component, (z,) y, x]\",\n" );
            std::fprintf( f, " \"blocks\": [\n" );
            const int* rp = cfg.ranks_per_dim;
            const int W = cfg.world_size;
            for ( int r = 0; r < W; ++r )
            {
                int b[3] = { r % rp[0], ( r / rp[0] ) % rp[1], D == 3 ? r / ( rp[0] * rp[1] ) : 0 };
                std::fprintf( f, "  {\"rank\": %d, \"offset\": [", r );
                int ext[3] = { 1, 1, 1 }, off[3] = { 0, 0, 0 };
                for ( int d = 0; d < D; ++d )
                    partition_of( cfg.global_num_cell[d], rp[d], b[d], ext[d], off[d] );
                for ( int d = 0; d < D; ++d )
                    std::fprintf( f, "%d%s", off[d], d + 1 < D ? ", " : "], \"extent\": [" );
                for ( int d = 0; d < D; ++d )
                    std::fprintf( f, "%d%s", ext[d], d + 1 < D ? ", " : "], " );
                std::fprintf( f, "\"quantity\": \"%s\", \"velocity\": \"%s\"}%s\n",
                              block_name( r, p.step, "quantity" ).c_str(), block_name( r, p.step, "velocity" ).c_str(),
                              r + 1 < W ? "," : "" );
            }
            std::fprintf( f, " ]}\n" );
            ok = std::fclose( f ) == 0;
        }
    }
    if ( !ok )
        return cfb_fail( c, CFB_ERR_INVALID, "writing output under " + p.dir + " failed: " + std::strerror( errno ) );
    return CFB_OK;
}

} // namespace

int output_flush( cfb_ctx* c )
{
    OutputStage* o = c->out;
    if ( !o )
        return CFB_OK;
    int rc = wait_copy( c );
    if ( rc )
        return rc;
    if ( o->pending.active )
    {
        o->pending.active = false;
        return write_files( c );
    }
    return CFB_OK;
}

int output_write( cfb_ctx* c, const char* dir, int time_step )
{
    int rc = stage_init( c );
    if ( rc )
        return rc;
    rc = output_flush( c ); // the previous write's files; frees the staging buffers for this one
    if ( rc )
        return rc;
    rc = extract_async( c );
    if ( rc )
        return rc;
    Pending& p = c->out->pending;
    p.active = true;
    p.dir = ( dir && dir[0] ) ? dir : "data";
    p.step = time_step;
    p.time = c->g.time;
    p.dt = c->g.dt;
    return CFB_OK;
}

const char* output_solve_dir( const cfb_ctx* c )
{
    return ( c->out && !c->out->solve_dir.empty() ) ? c->out->solve_dir.c_str() : nullptr;
}

void output_destroy( cfb_ctx* c )
{
    OutputStage* o = c->out;
    if ( !o )
        return;
    output_flush( c );
    if ( o->d_buf )
        cudaFree( o->d_buf );
    if ( o->h_buf )
        cudaFreeHost( o->h_buf );
    if ( o->ev_extracted )
        cudaEventDestroy( o->ev_extracted );
    if ( o->ev_copied )
        cudaEventDestroy( o->ev_copied );
    if ( o->io_stream )
        cudaStreamDestroy( o->io_stream );
    delete o;
    c->out = nullptr;
}

extern "C" {

int cfb_output_extract( cfb_ctx* c, double* quantity, double* velocity, double* nodes_x, double* nodes_y,
                        double* nodes_z )
{
    int rc = stage_init( c );
    if ( rc )
        return rc;
    rc = output_flush( c );
    if ( rc )
        return rc;
    rc = extract_async( c );
    if ( rc )
        return rc;
    rc = wait_copy( c );
    if ( rc )
        return rc;
    const Geo& g = c->g;
    const size_t ncell = ncell_of( g );
    if ( quantity )
        std::memcpy( quantity, c->out->h_buf, ncell * sizeof( double ) );
    if ( velocity )
        std::memcpy( velocity, c->out->h_buf + ncell, ncell * g.D * sizeof( double ) );
    double* nodes[3] = { nodes_x, nodes_y, nodes_z };
    for ( int d = 0; d < g.D; ++d )
        if ( nodes[d] )
            node_coordinates( g, d, nodes[d] );
    return CFB_OK;
}

int cfb_write_output( cfb_ctx* c, const char* dir, int time_step ) { return output_write( c, dir, time_step ); }

int cfb_write_npy( const char* path, const double* data, int ndim, const int64_t* shape )
{
    if ( !path || !data || ndim < 1 || ndim > 8 || !shape )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "write_npy: bad argument" );
    std::vector<long long> shp( shape, shape + ndim );
    if ( !write_npy( path, data, shp ) )
        return cfb_fail( nullptr, CFB_ERR_INVALID, std::string( "write_npy: " ) + path + ": " + std::strerror( errno ) );
    return CFB_OK;
}

int cfb_output_flush( cfb_ctx* c ) { return output_flush( c ); }

int cfb_set_output_dir( cfb_ctx* c, const char* dir )
{
    int rc = stage_init( c );
    if ( rc )
        return rc;
    c->out->solve_dir = dir ? dir : "";
    return CFB_OK;
}

} // extern "C"
