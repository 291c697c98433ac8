// host_tools.cu — host-side preprocessing that is too slow in Python at 1M links (plain C++, no device code).
#include <stdint.h>

#include <deque>
#include <vector>

#include "tarl_b200.h"

extern "C" {

// Locality ordering of the links for the resident link store: grows clusters of `cluster` links by breadth-first
// search over the (undirected view of the) dual graph and emits them one after the other, so that a CTA tile of
// `cluster` consecutive store slots holds links that are each other's upstream / downstream neighbours — their
// neighbour gathers then hit the lines the tile itself just loaded. The next cluster is seeded from the frontier the
// previous one left behind, which keeps consecutive tiles adjacent in the network (L2 locality).
// adj: CSR over links (host memory), each row listing the link's in- and out-neighbours. order[slot] = link id.
int tarl_cluster_links(int32_t n_links, const int32_t* adj_ptr, const int32_t* adj_idx, int32_t cluster,
                       int32_t* order) {
    if (n_links < 0 || cluster < 1 || (n_links > 0 && (!adj_ptr || !order))) return TARL_E_BADARG;
    std::vector<uint8_t> state(n_links, 0);   // 0 free, 1 queued in the current cluster, 2 placed
    std::deque<int32_t> seeds;                // frontier links left over by earlier clusters
    std::vector<int32_t> queue;
    int32_t placed = 0, scan = 0;
    while (placed < n_links) {
        int32_t seed = -1;
        while (!seeds.empty()) {
            const int32_t s = seeds.front();
            seeds.pop_front();
            if (state[s] == 0) { seed = s; break; }
        }
        if (seed < 0) {
            while (scan < n_links && state[scan] != 0) ++scan;
            seed = scan;
        }
        queue.clear();
        queue.push_back(seed);
        state[seed] = 1;
        size_t head = 0;
        int32_t count = 0;
        while (head < queue.size() && count < cluster) {
            const int32_t v = queue[head++];
            order[placed++] = v;
            state[v] = 2;
            ++count;
            for (int32_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k) {
                const int32_t w = adj_idx[k];
                if (w >= 0 && w < n_links && state[w] == 0) { state[w] = 1; queue.push_back(w); }
            }
        }
        for (; head < queue.size(); ++head) {   // not reached: free again, and candidates to seed the next cluster
            state[queue[head]] = 0;
            seeds.push_back(queue[head]);
        }
    }
    return TARL_OK;
}

}  // extern "C"
