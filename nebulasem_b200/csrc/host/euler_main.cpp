// euler_main.cpp -- `euler ./controls`: drop-in for the reference solver binary (apps/euler/euler.cpp:296-302).
// Reads the same controls / grid / field files from the current directory, runs the time loop on the GPU through
// the C ABI and writes the same rho/U/T/p<k>.{bin,txt} dumps.
//
// One process per GPU/partition, launched the way the reference is (`mpirun -np N euler ./controls`, test.sh) or by
// any launcher that exports a rank and a world size: this binary links no MPI, it reads the launcher's environment
// (OMPI_COMM_WORLD_RANK/SIZE, PMI_RANK/SIZE, SLURM_PROCID/NTASKS, torchrun's RANK/WORLD_SIZE, or NSEM_RANK/NSEM_WORLD).
// Every rank decomposes the global grid identically (decomposition{type n}, METIS by default) and keeps its part; the
// ncclUniqueId travels from rank 0 to the others through a file in the case directory (the reference hands its
// partitions over through the file system too, field.cpp:1086-1443); fields are dumped per rank into grid<r>/ and merged
// by rank 0 into the case directory in global node order (Prepare::mergeFields).
//   NSEM_DEVICE   device index (default: local rank, else rank % visible devices)
//   NSEM_DRYRUN=k (k >= 1) set-up only, no GPU: decompose, read and initialise the fields, dump them as index k
//   NSEM_VTK=1    also write <mesh><k>.vtk next to every dump (the file `prepare ./controls -vtk -start k` would make of it)
// `euler ./controls -vtk [-start i] [-stop j]` converts existing dumps i..j-1 of the case directory (the merged, global ones) to VTK on the
// host, no GPU involved: the post-processing step of the reference's workflow (apps/prepare/prepareApp.cpp:26-27,118-124) for this solver's fields.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "nsem_host.h"

static bool env_int(const char* name, int& out) {
    const char* v = std::getenv(name);
    if (!v || !*v) return false;
    out = std::atoi(v);
    return true;
}

// Launch handshake through the case directory (the reference hands its partitions over through the file system too): rank 0 hands every
// other rank the 128-byte ncclUniqueId and a 64-bit launch nonce (it stamps the per-dump markers rank 0's merge waits for).  Nothing here
// trusts a file's age: every rank r > 0 draws a fresh random nonce and keeps a hello file `.nsem_hello.<r>` alive; rank 0 removes whatever a
// crashed earlier launch left behind BEFORE it does anything else, publishes `.nsem_nccl_id` = blob + the nonces it has seen, and a rank
// accepts that file only if it carries its own nonce -- a stale file cannot.  Rank 0 returns once every rank has acknowledged.
static uint64_t random_u64() {
    uint64_t v = 0;
    FILE* f = std::fopen("/dev/urandom", "rb");
    if (f) { if (std::fread(&v, sizeof v, 1, f) != 1) v = 0; std::fclose(f); }
    if (v == 0) v = ((uint64_t)::getpid() << 32) ^ (uint64_t)std::time(nullptr) ^ 0x9e3779b97f4a7c15ull;
    return v;
}
static bool read_exact(const std::string& path, void* buf, size_t n) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    const size_t got = std::fread(buf, 1, n, f);
    unsigned char extra;
    const bool more = std::fread(&extra, 1, 1, f) == 1;
    std::fclose(f);
    return got == n && !more;
}
static void write_atomic(const std::string& path, const void* buf, size_t n) {
    const std::string tmp = path + ".tmp" + std::to_string(::getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(buf, 1, n, f) != n) { if (f) std::fclose(f); throw nsemh::Error("cannot write " + tmp); }
    std::fclose(f);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw nsemh::Error("cannot publish " + path);
}
struct LaunchBlob { unsigned char id[128]; uint64_t nonce; };
static void share_launch_blob(const std::string& dir, int rank, int world, bool with_id, LaunchBlob& blob) {
    const std::string idf = dir + "/.nsem_nccl_id";
    auto hello = [&](int r) { return dir + "/.nsem_hello." + std::to_string(r); };
    auto ack = [&](int r) { return dir + "/.nsem_ack." + std::to_string(r); };
    const size_t fsize = sizeof(LaunchBlob) + 8 * (size_t)(world - 1);
    const int max_tries = 6000;                                   // <= 10 min
    if (rank == 0) {
        ::unlink(idf.c_str());
        for (int r = 1; r < world; r++) { ::unlink(hello(r).c_str()); ::unlink(ack(r).c_str()); }
        std::memset(&blob, 0, sizeof blob);
        if (with_id && nsem_get_unique_id(blob.id)) throw nsemh::Error(nsem_last_error(nullptr));
        blob.nonce = random_u64();
        std::vector<uint64_t> seen(world, 0), published(world, 0);
        bool have_published = false;
        for (int tries = 0; tries < max_tries; tries++) {
            bool all = true;
            for (int r = 1; r < world; r++) all &= read_exact(hello(r), &seen[r], 8) && seen[r] != 0;
            if (all && (!have_published || seen != published)) {
                std::vector<unsigned char> buf(fsize);
                std::memcpy(buf.data(), &blob, sizeof blob);
                std::memcpy(buf.data() + sizeof blob, seen.data() + 1, 8 * (size_t)(world - 1));
                write_atomic(idf, buf.data(), fsize);
                published = seen;
                have_published = true;
            }
            if (have_published) {
                bool acked = true;
                for (int r = 1; r < world; r++) { uint64_t a = 0; acked &= read_exact(ack(r), &a, 8) && a == published[r]; }
                if (acked) {
                    ::unlink(idf.c_str());
                    for (int r = 1; r < world; r++) { ::unlink(hello(r).c_str()); ::unlink(ack(r).c_str()); }
                    return;
                }
            }
            ::usleep(100000);
        }
        ::unlink(idf.c_str());
        throw nsemh::Error("rank 0: the other ranks never answered the launch handshake in " + dir);
    }
    const uint64_t mine = random_u64();
    std::vector<unsigned char> buf(fsize);
    for (int tries = 0; tries < max_tries; tries++) {
        uint64_t cur = 0;
        if (!(read_exact(hello(rank), &cur, 8) && cur == mine)) write_atomic(hello(rank), &mine, 8);      // rank 0's start-up sweep may remove it
        if (read_exact(idf, buf.data(), fsize)) {
            uint64_t slot = 0;
            std::memcpy(&slot, buf.data() + sizeof blob + 8 * (size_t)(rank - 1), 8);
            if (slot == mine) {
                std::memcpy(&blob, buf.data(), sizeof blob);
                write_atomic(ack(rank), &mine, 8);
                return;
            }
        }
        ::usleep(100000);
    }
    ::unlink(hello(rank).c_str());
    throw nsemh::Error("rank " + std::to_string(rank) + ": rank 0 never answered the launch handshake in " + dir);
}

int main(int argc, char* argv[]) {
    if (argc < 2 || !std::strcmp(argv[1], "-h")) {
        std::printf("Usage:\n  %s <inputfile> <Options>\nOptions:\n  -vtk        --  Convert data to VTK format\n  -start <i>  --  Start at time step <i>\n"
                    "  -stop <i>   --  Stop before time step <i>\n  -h          --  Display this message\n\n", argv[0]);
        return argc < 2 ? 1 : 0;
    }
    int rank = 0, world = 1, local = -1;
    try {
        std::unique_ptr<nsemh::EulerSolver> sp(new nsemh::EulerSolver());
        nsemh::EulerSolver& s = *sp;
        std::string ctl = argv[1];
        std::string dir = ".";
        const size_t slash = ctl.find_last_of('/');
        if (slash != std::string::npos && ctl.substr(slash + 1) == "controls") dir = ctl.substr(0, slash);
        else if (ctl != "controls" && ctl != "./controls") {
            std::fprintf(stderr, "euler: the input file must be named `controls`\n");
            return 1;
        }
        if (!(env_int("NSEM_RANK", rank) && env_int("NSEM_WORLD", world)) &&
            !(env_int("OMPI_COMM_WORLD_RANK", rank) && env_int("OMPI_COMM_WORLD_SIZE", world)) &&
            !(env_int("PMI_RANK", rank) && env_int("PMI_SIZE", world)) &&
            !(env_int("SLURM_PROCID", rank) && env_int("SLURM_NTASKS", world)) &&
            !(env_int("RANK", rank) && env_int("WORLD_SIZE", world))) { rank = 0; world = 1; }
        if (world < 1 || rank < 0 || rank >= world) throw nsemh::Error("inconsistent rank/world size in the environment");
        if (!env_int("NSEM_DEVICE", local) && !env_int("OMPI_COMM_WORLD_LOCAL_RANK", local) && !env_int("LOCAL_RANK", local) &&
            !env_int("SLURM_LOCALID", local)) local = -1;                     // -1: rank % visible devices (nsem_create)
        bool to_vtk = false;
        int vtk_start = 0, vtk_stop = 0;
        for (int i = 2; i < argc; i++) {
            if (!std::strcmp(argv[i], "-vtk")) to_vtk = true;
            else if (!std::strcmp(argv[i], "-start") && i + 1 < argc) { vtk_start = std::atoi(argv[++i]); vtk_stop = vtk_start + 1; }
            else if (!std::strcmp(argv[i], "-stop") && i + 1 < argc) vtk_stop = std::atoi(argv[++i]);
            else throw nsemh::Error(std::string("unknown option ") + argv[i]);
        }
        if (to_vtk) {
            if (rank != 0) return 0;                       // the merged dumps are global: one process converts them
            rank = 0; world = 1;
            s.vtk_mode = true;
            s.read_controls(dir);
            std::printf("Converting result to VTK format.\n");
            for (int k = vtk_start; k < vtk_stop; k++) {
                s.load_mesh(k);
                s.read_fields(k);
                s.write_vtk(k);
            }
            std::printf("Exiting application run with %d processes\n", world);
            return 0;
        }
        s.rank = rank; s.nranks = world;
        s.read_controls(dir);
        const int step = (int)s.start_step;       // dump index (read_controls divides by write_interval)
        s.load_mesh(step);
        s.read_fields(step);
        s.setup();
        const char* dry = std::getenv("NSEM_DRYRUN");
        const bool dryrun = dry && std::atoi(dry) > 0;
        LaunchBlob blob;
        std::memset(&blob, 0, sizeof blob);
        if (world > 1) share_launch_blob(dir, rank, world, !dryrun, blob);
        s.launch_nonce = blob.nonce;
        if (dryrun) {
            s.write_fields(std::atoi(dry));
            s.merge_fields(std::atoi(dry));
        } else {
            s.attach_device(world > 1 ? local : (local < 0 ? 0 : local), rank, world, world > 1 ? blob.id : nullptr);
            nsemh::run_case(sp);        // `s` is gone after a regrid: nothing below touches it
        }
        if (rank == 0) std::printf("Exiting application run with %d processes\n", world);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "euler[%d/%d]: %s\n", rank, world, e.what());
        return 1;
    }
}
