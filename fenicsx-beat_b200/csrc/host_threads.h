// Host-side thread helpers shared by the set-up routines (fem_assemble.cu, csr_patterns.cu).  No device code.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

// MONO_HOST_THREADS, default min(hardware threads, 16)
inline int host_threads() {
    if (const char* e = std::getenv("MONO_HOST_THREADS")) {
        int v = std::atoi(e);
        if (v > 0) return v;
    }
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
}

template <class F>
inline void run_threads(int nt, F&& body) {  // body(thread id)
    if (nt <= 1) {
        body(0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back([&body, t] { body(t); });
    for (auto& th : pool) th.join();
}

