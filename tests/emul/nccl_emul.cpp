// nccl_emul.cpp — tests/emul only: in-process NCCL stand-in (ranks are threads) + the dlopen / dlsym
// replacements through which halo.cu finds it.
#include "nccl.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace
{

struct Group
{
    int nranks = 0;
    std::mutex mu;
    std::condition_variable cv;
    // mailbox[src * nranks + dst]: buffered messages in order
    std::vector<std::deque<std::vector<char>>> mail;
    // collectives
    int arrived = 0;
    long long generation = 0;
    std::vector<std::vector<char>> contrib;
    std::vector<char> result;
};

std::mutex g_mu;
std::map<std::string, std::shared_ptr<Group>> g_groups;
long long g_next_id = 1;

size_t type_size( ncclDataType_t t ) { return t == ncclDouble ? 8 : ( t == ncclInt ? 4 : 1 ); }

struct PendingOp
{
    bool send;
    const void* sbuf;
    void* rbuf;
    size_t bytes;
    int peer;
    struct cfb_emul_comm* comm;
};
thread_local int t_group_depth = 0;
thread_local std::vector<PendingOp> t_pending;

} // namespace

struct cfb_emul_comm
{
    std::shared_ptr<Group> g;
    int rank = 0;
};

namespace
{

void do_send( cfb_emul_comm* c, const void* buf, size_t bytes, int peer )
{
    Group& g = *c->g;
    std::vector<char> m( bytes );
    std::memcpy( m.data(), buf, bytes );
    {
        std::lock_guard<std::mutex> lk( g.mu );
        g.mail[(size_t)c->rank * g.nranks + peer].push_back( std::move( m ) );
    }
    g.cv.notify_all();
}

void do_recv( cfb_emul_comm* c, void* buf, size_t bytes, int peer )
{
    Group& g = *c->g;
    std::unique_lock<std::mutex> lk( g.mu );
    auto& q = g.mail[(size_t)peer * g.nranks + c->rank];
    g.cv.wait( lk, [&]() { return !q.empty(); } );
    std::memcpy( buf, q.front().data(), bytes < q.front().size() ? bytes : q.front().size() );
    q.pop_front();
}

// every rank contributes `bytes`, every rank receives all contributions in rank order.  The result of one
// collective is replaced only when ALL ranks have arrived at the next one, which a rank cannot do before
// it has copied this one out.
void all_gather( cfb_emul_comm* c, const void* send, void* recv, size_t bytes )
{
    Group& g = *c->g;
    std::unique_lock<std::mutex> lk( g.mu );
    const long long gen = g.generation;
    g.contrib[c->rank].assign( static_cast<const char*>( send ), static_cast<const char*>( send ) + bytes );
    if ( ++g.arrived == g.nranks )
    {
        g.result.clear();
        for ( int r = 0; r < g.nranks; ++r )
            g.result.insert( g.result.end(), g.contrib[r].begin(), g.contrib[r].end() );
        g.arrived = 0;
        ++g.generation;
        g.cv.notify_all();
    }
    else
        g.cv.wait( lk, [&]() { return g.generation != gen; } );
    std::vector<char> mine = g.result;
    lk.unlock();
    std::memcpy( recv, mine.data(), mine.size() );
}

ncclResult_t GetUniqueId( ncclUniqueId* id )
{
    std::lock_guard<std::mutex> lk( g_mu );
    std::memset( id, 0, sizeof( *id ) );
    std::snprintf( id->internal, sizeof( id->internal ), "cfb-emul-%lld", g_next_id++ );
    return ncclSuccess;
}
ncclResult_t CommInitRank( ncclComm_t* comm, int nranks, ncclUniqueId id, int rank )
{
    std::lock_guard<std::mutex> lk( g_mu );
    auto& g = g_groups[std::string( id.internal )];
    if ( !g )
    {
        g = std::make_shared<Group>();
        g->nranks = nranks;
        g->mail.resize( (size_t)nranks * nranks );
        g->contrib.resize( nranks );
    }
    if ( g->nranks != nranks || rank < 0 || rank >= nranks )
        return ncclInvalidArgument;
    *comm = new cfb_emul_comm{ g, rank };
    return ncclSuccess;
}
ncclResult_t CommDestroy( ncclComm_t c )
{
    delete c;
    return ncclSuccess;
}
ncclResult_t GroupStart()
{
    ++t_group_depth;
    return ncclSuccess;
}
ncclResult_t GroupEnd()
{
    if ( --t_group_depth > 0 )
        return ncclSuccess;
    // sends are buffered, so posting all of them first can never deadlock
    for ( auto& op : t_pending )
        if ( op.send )
            do_send( op.comm, op.sbuf, op.bytes, op.peer );
    for ( auto& op : t_pending )
        if ( !op.send )
            do_recv( op.comm, op.rbuf, op.bytes, op.peer );
    t_pending.clear();
    return ncclSuccess;
}
ncclResult_t Send( const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t )
{
    if ( t_group_depth > 0 )
        t_pending.push_back( { true, buf, nullptr, count * type_size( t ), peer, c } );
    else
        do_send( c, buf, count * type_size( t ), peer );
    return ncclSuccess;
}
ncclResult_t Recv( void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t )
{
    if ( t_group_depth > 0 )
        t_pending.push_back( { false, nullptr, buf, count * type_size( t ), peer, c } );
    else
        do_recv( c, buf, count * type_size( t ), peer );
    return ncclSuccess;
}
ncclResult_t AllGather( const void* send, void* recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t )
{
    all_gather( c, send, recv, count * type_size( t ) );
    return ncclSuccess;
}
ncclResult_t AllReduce( const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t, ncclComm_t c,
                        cudaStream_t )
{
    if ( t != ncclDouble )
        return ncclInvalidArgument;
    const int n = c->g->nranks;
    std::vector<double> all( count * n );
    all_gather( c, send, all.data(), count * sizeof( double ) );
    double* out = static_cast<double*>( recv );
    for ( size_t i = 0; i < count; ++i )
    {
        double s = 0.0;
        for ( int r = 0; r < n; ++r )
            s += all[(size_t)r * count + i];
        out[i] = s;
    }
    return ncclSuccess;
}
const char* GetErrorString( ncclResult_t ) { return "emulated NCCL error"; }

struct Sym
{
    const char* name;
    void* fn;
};
const Sym g_syms[] = { { "ncclGetUniqueId", (void*)GetUniqueId },   { "ncclCommInitRank", (void*)CommInitRank },
                       { "ncclCommDestroy", (void*)CommDestroy },   { "ncclSend", (void*)Send },
                       { "ncclRecv", (void*)Recv },                 { "ncclAllReduce", (void*)AllReduce },
                       { "ncclAllGather", (void*)AllGather },       { "ncclGroupStart", (void*)GroupStart },
                       { "ncclGroupEnd", (void*)GroupEnd },         { "ncclGetErrorString", (void*)GetErrorString } };

} // namespace

// halo.cu is compiled with -Ddlopen=cfb_emul_dlopen -Ddlsym=cfb_emul_dlsym -Ddlerror=cfb_emul_dlerror
extern "C" void* cfb_emul_dlopen( const char*, int )
{
    static int handle;
    return &handle;
}
extern "C" void* cfb_emul_dlsym( void*, const char* name )
{
    for ( const Sym& s : g_syms )
        if ( std::strcmp( s.name, name ) == 0 )
            return s.fn;
    return nullptr;
}
extern "C" char* cfb_emul_dlerror( void )
{
    static char msg[] = "emulated dlerror";
    return msg;
}
