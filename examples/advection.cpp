// advection.cpp — drop-in counterpart of the reference's examples/advection.cpp for the B200 path.
//
// Same command line (short and long options, defaults, and quirks: -w is the inflow WIDTH, -h the
// inflow HEIGHT, --help only in long form, -i is the timestep, -p the driver; reference:
// examples/advection.cpp:43-68,160-379), same problem set-up (all walls SOLID, zero initial state,
// gravity along -y; :438-459), but the solver is created with the device string "b200" and the
// matrix solver defaults to "Reference" — the only solver this backend implements (no HYPRE).
// Extra: -D 2|3 selects the space dimension (the reference is 2-D only), -z/-e/--input-velocity-z
// extend the inflow box in z, -o FILE dumps the final q field as raw float64 for comparisons.
//
// Multi-GPU: launch one process per GPU with CFB_RANK / CFB_WORLD_SIZE in the environment and a
// shared file (CFB_NCCL_ID_FILE) that rank 0 fills with the NCCL ids (what MPI_Bcast would do).
#include <cajitafluids_b200/CajitaFluids.hpp>

#include <getopt.h>

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <thread>

namespace
{

struct ClArgs
{
    std::string device = "b200";
    int dim = 2;
    int cells = 128;
    double size = 1.0;
    double t_final = 4.0;
    double delta_t = 0.005;
    int write_freq = 20;
    double density = 0.1;
    double gravity = 0.0;
    std::string solver = "Reference";
    std::string precon = "None";
    double in_loc[3] = { 0.2, 0.45, 0.45 };
    double in_size[3] = { 0.02, 0.1, 0.1 };
    double in_vel[3] = { 1.0, 0.0, 0.0 };
    double in_quantity = 3.0;
    std::string dump;
    std::string out_dir = "data"; // the reference always writes into data/ (src/SiloWriter.hpp:379-384)
};

const char* short_opts = "n:s:t:i:d:g:p:m:c:x:y:z:w:h:e:q:u:v:D:o:";
const option long_opts[] = { { "cells", required_argument, nullptr, 'n' },
                             { "size", required_argument, nullptr, 's' },
                             { "time", required_argument, nullptr, 't' },
                             { "deltat", required_argument, nullptr, 'i' },
                             { "density", required_argument, nullptr, 'd' },
                             { "gravity", required_argument, nullptr, 'g' },
                             { "driver", required_argument, nullptr, 'p' },
                             { "matrix-solver", required_argument, nullptr, 'm' },
                             { "preconditioner", required_argument, nullptr, 'c' },
                             { "input-x", required_argument, nullptr, 'x' },
                             { "input-y", required_argument, nullptr, 'y' },
                             { "input-z", required_argument, nullptr, 'z' },
                             { "input-width", required_argument, nullptr, 'w' },
                             { "input-height", required_argument, nullptr, 'h' },
                             { "input-depth", required_argument, nullptr, 'e' },
                             { "input-quantity", required_argument, nullptr, 'q' },
                             { "input-velocity-x", required_argument, nullptr, 'u' },
                             { "input-velocity-y", required_argument, nullptr, 'v' },
                             { "input-velocity-z", required_argument, nullptr, 1001 },
                             { "dim", required_argument, nullptr, 'D' },
                             { "dump", required_argument, nullptr, 'o' },
                             { "output-dir", required_argument, nullptr, 1002 },
                             { "help", no_argument, nullptr, 'j' },
                             { nullptr, 0, nullptr, 0 } };

void usage( const char* prog )
{
    std::cerr << "Usage: " << prog << " [options]\n"
              << "  -n, --cells N            cells per side (default 128)\n"
              << "  -s, --size L             domain edge length (default 1.0)\n"
              << "  -t, --time T             simulated time (default 4.0)\n"
              << "  -i, --deltat DT          timestep (default 0.005, clamped to h/umax)\n"
              << "  -d, --density RHO        fluid density (default 0.1)\n"
              << "  -g, --gravity G          gravity along -y (default 0)\n"
              << "  -p, --driver NAME        b200 (only backend)\n"
              << "  -m, --matrix-solver S    Reference (only solver; HYPRE names are rejected)\n"
              << "  -c, --preconditioner P   None/Jacobi/Diagonal: the Reference solver's built-in Jacobi (default);\n"
              << "                           MG: opt-in geometric multigrid V(2,2) cycle (same system, same\n"
              << "                           stopping test, O(10) CG iterations instead of O(n))\n"
              << "  -x/-y/-z, -w/-h/-e       inflow box corner and extent\n"
              << "  -q, -u, -v               inflow quantity and velocity\n"
              << "  -D, --dim 2|3            space dimension (default 2, like the reference)\n"
              << "  -o, --dump FILE          write the final q (owned cells, float64) to FILE\n"
              << "      --output-dir DIR     where the periodic output goes (default data, like the reference;\n"
              << "                           .npy per variable and block + a .json master per written step;\n"
              << "                           'none' turns it off)\n";
}

double positive( const char* what, const char* arg )
{
    double v = std::atof( arg );
    if ( v <= 0.0 )
    {
        std::cerr << "Invalid " << what << " argument.\n";
        std::exit( -1 );
    }
    return v;
}

int parse( int argc, char** argv, ClArgs& cl )
{
    int ch;
    while ( ( ch = getopt_long( argc, argv, short_opts, long_opts, nullptr ) ) != -1 )
    {
        switch ( ch )
        {
        case 'n':
            cl.cells = (int)positive( "cells", optarg );
            break;
        case 's':
            cl.size = positive( "size", optarg );
            break;
        case 't':
            cl.t_final = positive( "timesteps", optarg );
            break;
        case 'i':
            cl.delta_t = positive( "timestep", optarg );
            break;
        case 'd':
            cl.density = positive( "density", optarg );
            break;
        case 'g':
            cl.gravity = std::atof( optarg );
            break;
        case 'p':
            cl.device = optarg;
            break;
        case 'm':
            cl.solver = optarg;
            break;
        case 'c':
            cl.precon = optarg;
            break;
        case 'x':
            cl.in_loc[0] = std::atof( optarg );
            break;
        case 'y':
            cl.in_loc[1] = std::atof( optarg );
            break;
        case 'z':
            cl.in_loc[2] = std::atof( optarg );
            break;
        case 'w':
            cl.in_size[0] = positive( "inflow width", optarg );
            break;
        case 'h':
            cl.in_size[1] = positive( "inflow height", optarg );
            break;
        case 'e':
            cl.in_size[2] = positive( "inflow depth", optarg );
            break;
        case 'q':
            cl.in_quantity = std::atof( optarg );
            break;
        case 'u':
            cl.in_vel[0] = std::atof( optarg );
            break;
        case 'v':
            cl.in_vel[1] = std::atof( optarg );
            break;
        case 1001:
            cl.in_vel[2] = std::atof( optarg );
            break;
        case 'D':
            cl.dim = std::atoi( optarg );
            break;
        case 'o':
            cl.dump = optarg;
            break;
        case 1002:
            cl.out_dir = std::strcmp( optarg, "none" ) == 0 ? "" : optarg;
            break;
        case 'j':
            usage( argv[0] );
            std::exit( 0 );
        default:
            usage( argv[0] );
            std::exit( -1 );
        }
    }
    if ( cl.dim != 2 && cl.dim != 3 )
    {
        std::cerr << "Invalid dimension argument.\n";
        std::exit( -1 );
    }
    return 0;
}

// Initial state: constant quantity and velocity (the reference's MeshInitFunc)
template <std::size_t Dim>
struct MeshInitFunc
{
    double _q;
    std::array<double, Dim> _u;
    template <class Entity>
    bool operator()( Entity, CajitaFluids::Field::Quantity, const int*, const double*, double& quantity ) const
    {
        quantity = _q;
        return true;
    }
    template <class Entity>
    bool operator()( Entity, CajitaFluids::Field::Velocity, const int*, const double*, double& velocity ) const
    {
        velocity = _u[Entity::id - 1];
        return true;
    }
};

CajitaFluids::Comm make_comm()
{
    CajitaFluids::Comm comm;
    const char* r = std::getenv( "CFB_RANK" );
    const char* s = std::getenv( "CFB_WORLD_SIZE" );
    comm.rank = r ? std::atoi( r ) : 0;
    comm.size = s ? std::atoi( s ) : 1;
    if ( comm.size > 1 )
    {
        const char* f = std::getenv( "CFB_NCCL_ID_FILE" );
        if ( !f )
            throw std::runtime_error( "CFB_NCCL_ID_FILE must name a shared file for the NCCL ids" );
        if ( comm.rank == 0 )
        {
            if ( cfb_nccl_unique_id( comm.nccl_id.data() ) != CFB_OK )
                throw std::runtime_error( cfb_last_error( nullptr ) );
            std::ofstream out( std::string( f ) + ".tmp", std::ios::binary );
            out.write( reinterpret_cast<const char*>( comm.nccl_id.data() ), comm.nccl_id.size() );
            out.close();
            std::rename( ( std::string( f ) + ".tmp" ).c_str(), f );
        }
        else
        {
            for ( int tries = 0;; ++tries )
            {
                std::ifstream in( f, std::ios::binary );
                if ( in && in.read( reinterpret_cast<char*>( comm.nccl_id.data() ), comm.nccl_id.size() ) )
                    break;
                if ( tries > 600 )
                    throw std::runtime_error( "timed out waiting for the NCCL id file" );
                std::this_thread::sleep_for( std::chrono::milliseconds( 100 ) );
            }
        }
    }
    return comm;
}

template <std::size_t Dim>
int advect( const ClArgs& cl )
{
    using namespace CajitaFluids;
    Comm comm = make_comm();
    DimBlockPartitioner<Dim> partitioner;
    BoundaryCondition<Dim> bc;
    bc.boundary_type.fill( BoundaryType::SOLID );

    std::array<double, Dim> loc, size, vel;
    std::array<double, 2 * Dim> box;
    std::array<int, Dim> ncell;
    for ( std::size_t d = 0; d < Dim; ++d )
    {
        loc[d] = cl.in_loc[d];
        size[d] = cl.in_size[d];
        vel[d] = cl.in_vel[d];
        box[d] = 0.0;
        box[Dim + d] = cl.size;
        ncell[d] = cl.cells;
    }
    InflowSource<Dim> source( loc, size, vel, cl.in_quantity );
    BodyForce<Dim> body( 0.0, -cl.gravity );
    MeshInitFunc<Dim> initializer{ 0.0, {} };

    auto solver = createSolver<Dim>( cl.device, comm, box, ncell, partitioner, cl.density, initializer, bc, source,
                                     body, cl.delta_t, cl.solver, cl.precon );
    if ( auto* sv = dynamic_cast<Solver<Dim>*>( solver.get() ) )
        sv->siloWriter()->setDirectory( cl.out_dir );
    auto t0 = std::chrono::steady_clock::now();
    solver->solve( cl.t_final, cl.write_freq );
    double sec = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();

    auto* s = dynamic_cast<Solver<Dim>*>( solver.get() );
    if ( comm.rank == 0 && s )
    {
        cfb_stats st = s->stats();
        std::cout << "Finished " << s->stepsTaken() << " steps (t = " << s->time() << ") in " << sec << " s: "
                  << s->stepsTaken() / sec << " steps/s, " << st.cg_iterations << " CG iterations, "
                  << st.kernel_launches << " kernel launches\n"
                  << "  device ms: advect " << st.ms_advect << ", inputs " << st.ms_add_inputs << ", rhs "
                  << st.ms_build_rhs << ", pcg " << st.ms_pcg << ", apply " << st.ms_apply_pressure << "\n";
    }
    if ( !cl.dump.empty() && s )
    {
        auto q = s->problemManager()->copyToHost( Cell(), Version::Current() );
        std::ofstream out( cl.dump + ( comm.size > 1 ? "." + std::to_string( comm.rank ) : "" ), std::ios::binary );
        out.write( reinterpret_cast<const char*>( q.data() ), q.size() * sizeof( double ) );
    }
    return 0;
}

} // namespace

int main( int argc, char* argv[] )
{
    ClArgs cl;
    parse( argc, argv, cl );
    const char* r = std::getenv( "CFB_RANK" );
    if ( !r || std::atoi( r ) == 0 )
    {
        std::cout << "CajitaFluids (b200 backend)\n"
                  << "=======Command line arguments=======\n"
                  << std::left << std::setw( 22 ) << "Driver" << ": " << cl.device << "\n"
                  << std::setw( 22 ) << "Dimension" << ": " << cl.dim << "\n"
                  << std::setw( 22 ) << "Cells" << ": " << cl.cells << "\n"
                  << std::setw( 22 ) << "Domain" << ": " << cl.size << "\n"
                  << std::setw( 22 ) << "Input Flow" << ": " << cl.in_quantity << " at (" << cl.in_loc[0] << ", "
                  << cl.in_loc[1] << ") size (" << cl.in_size[0] << ", " << cl.in_size[1] << ") velocity ("
                  << cl.in_vel[0] << ", " << cl.in_vel[1] << ")\n"
                  << std::setw( 22 ) << "Total Simulation Time" << ": " << cl.t_final << "\n"
                  << std::setw( 22 ) << "Timestep Size" << ": " << cl.delta_t << "\n"
                  << std::setw( 22 ) << "Write Frequency" << ": " << cl.write_freq << "\n"
                  << std::setw( 22 ) << "Matrix Solver" << ": " << cl.solver << "\n"
                  << "====================================\n";
    }
    try
    {
        return cl.dim == 2 ? advect<2>( cl ) : advect<3>( cl );
    }
    catch ( const std::exception& e )
    {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
