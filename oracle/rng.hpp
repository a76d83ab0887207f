// ORACLE (test infrastructure, NOT product code).
//
// Two random-number back-ends for the restated hot path:
//
//  * DsfmtRng  - the reference's own stream: dSFMT 2.2.3 (MEXP 19937) compiled from the
//    reference tree into oracle/_ref/libhande_ref_c.so, consumed through a 50,000-double
//    close-open refill buffer exactly like lib/dSFMT_F03_interface/dSFMT_interface.F90:192-250,
//    387-409 (seed = rng_seed + iproc, src/fciqmc.f90:206).  Used to reproduce the golden
//    benchmark tables of test_suite/.
//
//  * PhiloxRng - the counter-based stream the B200 engine uses: Philox4x32-10 (Salmon et al.,
//    SC'11; public algorithm) keyed by (seed, MC cycle) with counter
//    (purpose<<24 | draw/2, attempt, hash64(determinant)).  A stochastic decision therefore
//    depends only on (seed, cycle, determinant, attempt, draw#) - not on list order, thread
//    scheduling or the number of ranks - so kernel and oracle agree bit for bit.
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include <stdexcept>
#include <dlfcn.h>
#include "system.hpp"

namespace oracle {

enum RngPurpose : uint32_t {
    RNG_NATTEMPTS = 0,   // decide_nattempts
    RNG_SPAWN = 1,       // excitation generator + attempt_to_spawn (per attempt)
    RNG_DEATH = 2,       // stochastic_death
    RNG_ROUND_MAIN = 3,  // remove_unoccupied_dets / stochastic_round
    RNG_ROUND_SPAWN = 4, // round_low_population_spawns
    RNG_DETERM = 5,      // deterministic_annihilation / create_spawned_particle_determ (semi-stochastic)
};

struct Rng {
    virtual ~Rng() {}
    // Select the stream for the next draws (no-op for sequential generators).
    virtual void begin(uint32_t purpose, const Det& f, int W, uint32_t attempt) = 0;
    virtual void set_cycle(uint32_t cycle) = 0;
    virtual double next() = 0;  // uniform in [0,1)
    // Add a salt to the stream selector chosen by begin() (counter-based generators only): CCMC streams are keyed by
    // (attempt, rank) instead of a determinant.
    virtual void mix(uint64_t salt) { (void)salt; }
    uint64_t ndraws = 0;
};

// ----------------------------------------------------------------------------- dSFMT
struct DsfmtLib {
    void* h = nullptr;
    void* (*malloc_dsfmt_t)() = nullptr;
    void (*free_dsfmt_t)(void*) = nullptr;
    void (*chk_init_gen_rand)(void*, uint32_t, int) = nullptr;
    void (*fill_array_close_open)(void*, double*, int) = nullptr;
    int (*get_min_array_size)() = nullptr;
    uint32_t (*murmur2)(const void*, int, uint32_t) = nullptr;
    static DsfmtLib& get(const std::string& path = "") {
        static DsfmtLib lib;
        if (!lib.h) {
            if (path.empty()) throw std::runtime_error("oracle/_ref/libhande_ref_c.so path not set");
            lib.h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!lib.h) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
            lib.malloc_dsfmt_t = (void* (*)())dlsym(lib.h, "malloc_dsfmt_t");
            lib.free_dsfmt_t = (void (*)(void*))dlsym(lib.h, "free_dsfmt_t");
            lib.chk_init_gen_rand = (void (*)(void*, uint32_t, int))dlsym(lib.h, "dsfmt_chk_init_gen_rand");
            lib.fill_array_close_open = (void (*)(void*, double*, int))dlsym(lib.h, "dsfmt_fill_array_close_open");
            lib.get_min_array_size = (int (*)())dlsym(lib.h, "dsfmt_get_min_array_size");
            lib.murmur2 = (uint32_t (*)(const void*, int, uint32_t))dlsym(lib.h, "MurmurHash2");
            if (!lib.malloc_dsfmt_t || !lib.chk_init_gen_rand || !lib.fill_array_close_open)
                throw std::runtime_error("missing dSFMT symbols in reference C library");
        }
        return lib;
    }
};

struct DsfmtRng : Rng {
    void* state = nullptr;
    std::vector<double> store;
    int next_element = 0;  // 0-based index; == size => refill
    DsfmtRng(int seed, int store_size = 50000) {
        DsfmtLib& L = DsfmtLib::get();
        int n = std::max(store_size, L.get_min_array_size());
        store.resize(n);
        next_element = n;
        state = L.malloc_dsfmt_t();
        L.chk_init_gen_rand(state, (uint32_t)seed, 19937);
    }
    ~DsfmtRng() override { if (state) DsfmtLib::get().free_dsfmt_t(state); }
    void begin(uint32_t, const Det&, int, uint32_t) override {}
    void set_cycle(uint32_t) override {}
    double next() override {
        if (next_element == (int)store.size()) {
            DsfmtLib::get().fill_array_close_open(state, store.data(), (int)store.size());
            next_element = 0;
        }
        ndraws++;
        return store[next_element++];
    }
};

// ----------------------------------------------------------------------------- Philox
struct Philox4x32 {
    static inline void round(uint32_t* c, uint32_t k0, uint32_t k1) {
        const uint64_t M0 = 0xD2511F53ull, M1 = 0xCD9E8D57ull;
        uint64_t p0 = M0 * c[0], p1 = M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void gen(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; ++r) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// 64-bit mix of the determinant words (splitmix64 finaliser chained over the words).
inline uint64_t det_hash64(const Det& f, int W) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < W; ++i) {
        uint64_t z = f.w[i] + h;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        h = z + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    }
    return h;
}

inline double u64_to_unit_double(uint32_t lo, uint32_t hi) {
    uint64_t u = ((uint64_t)hi << 32) | lo;
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);  // 2^-53
}

struct PhiloxRng : Rng {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t purpose = 0, draw = 0;
    uint32_t buf[4] = {0, 0, 0, 0};
    explicit PhiloxRng(uint32_t seed) { key[0] = seed; key[1] = 0; ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0; }
    void set_cycle(uint32_t cycle) override { key[1] = cycle; }
    void begin(uint32_t purpose_, const Det& f, int W, uint32_t attempt) override {
        uint64_t h = det_hash64(f, W);
        purpose = purpose_;
        ctr[1] = attempt;
        ctr[2] = (uint32_t)h;
        ctr[3] = (uint32_t)(h >> 32);
        draw = 0;
    }
    void mix(uint64_t salt) override {
        uint64_t h = ((uint64_t)ctr[3] << 32) | ctr[2];
        h += salt;
        ctr[2] = (uint32_t)h;
        ctr[3] = (uint32_t)(h >> 32);
    }
    double next() override {
        if ((draw & 1u) == 0) {
            ctr[0] = (purpose << 24) | (draw >> 1);
            Philox4x32::gen(ctr, key, buf);
        }
        double r = (draw & 1u) ? u64_to_unit_double(buf[2], buf[3]) : u64_to_unit_double(buf[0], buf[1]);
        draw++;
        ndraws++;
        return r;
    }
};

}  // namespace oracle
