// key_cache.cu -- per-key precomputations of the single-kernel gadget products, kept across calls for keys the caller has PINNED.
//
// The gadget kernels do not read a prepared key as vmp_prepare left it: the NTT120 kernel wants the collapsed key and the bit bound of the
// key's integer coefficients (ntt120_gadget.cu: three pre-pass launches per call), the FFT64 kernel a re-laid-out copy (fft64_gadget.cu:
// one pre-pass).  They are functions of the key alone, but a silent cache would be wrong the moment the caller rewrites the key's memory
// (every buffer is caller-owned), so caching is opt-in: pgb_gadget_key_pin(key) promises that the bytes of `key` stay unchanged until
// pgb_gadget_key_unpin(key) -- or until pgb_vmp_prepare writes that key again, which drops its cached forms.  An unpinned key is
// re-derived on every call, exactly as before.
#include <vector>

#include "internal.h"

struct KeyCacheEntry {
    const char *key;
    uint64_t key_len;
    uint64_t sig[KEY_SIG_WORDS];
    void *dev;       // device-resident cached form (may be null for host-only entries)
    int64_t host_val; // host-resident cached scalar (e.g. the bit bound of the key's coefficients); 0 = not set
};
struct KeyCache {
    std::vector<std::pair<const char *, uint64_t>> pinned; // (data pointer, bytes)
    std::vector<KeyCacheEntry> entries;
};

static KeyCache *cache_of(pgb_module *m, bool create) {
    if (!m->key_cache && create) m->key_cache = new KeyCache();
    return m->key_cache;
}

bool key_is_pinned(const pgb_module *m, const void *key) {
    const KeyCache *c = m->key_cache;
    if (!c) return false;
    for (const auto &p : c->pinned)
        if (p.first == (const char *)key) return true;
    return false;
}

void *key_cache_find(pgb_module *m, const void *key, const uint64_t *sig) {
    KeyCache *c = m->key_cache;
    if (!c) return nullptr;
    for (const auto &e : c->entries)
        if (e.key == (const char *)key && memcmp(e.sig, sig, sizeof e.sig) == 0) return e.dev;
    return nullptr;
}

int key_cache_insert(pgb_module *m, const void *key, uint64_t key_len, const uint64_t *sig, size_t bytes, void **out) {
    KeyCache *c = cache_of(m, true);
    KeyCacheEntry e;
    e.key = (const char *)key;
    e.key_len = key_len;
    memcpy(e.sig, sig, sizeof e.sig);
    e.dev = nullptr;
    e.host_val = 0;
    PGB_CHECK_CUDA(cudaSetDevice(m->device));
    PGB_CHECK_CUDA(cudaMalloc(&e.dev, bytes));
    c->entries.push_back(e);
    *out = e.dev;
    return PGB_OK;
}

// host-only entry of a pinned key (created on first use): a scalar derived from the key that the HOST needs before a launch
int64_t *key_cache_host_slot(pgb_module *m, const void *key, uint64_t key_len, const uint64_t *sig) {
    KeyCache *c = cache_of(m, true);
    for (auto &e : c->entries)
        if (e.key == (const char *)key && memcmp(e.sig, sig, sizeof e.sig) == 0) return &e.host_val;
    KeyCacheEntry e;
    e.key = (const char *)key;
    e.key_len = key_len;
    memcpy(e.sig, sig, sizeof e.sig);
    e.dev = nullptr;
    e.host_val = 0;
    c->entries.push_back(e);
    return &c->entries.back().host_val;
}

// drops every cached form of a key whose bytes overlap [p, p + len)
void key_cache_invalidate(pgb_module *m, const void *p, uint64_t len) {
    KeyCache *c = m->key_cache;
    if (!c || c->entries.empty()) return;
    bool synced = false;
    for (size_t i = 0; i < c->entries.size();) {
        const KeyCacheEntry &e = c->entries[i];
        if (ranges_overlap(e.key, e.key_len, p, len)) {
            if (!synced) {
                cudaStreamSynchronize(m->stream); // a kernel reading the cached form may still be in flight
                synced = true;
            }
            if (e.dev) cudaFree(e.dev);
            c->entries.erase(c->entries.begin() + (long)i);
        } else {
            i++;
        }
    }
}

void key_cache_destroy(pgb_module *m) {
    KeyCache *c = m->key_cache;
    if (!c) return;
    for (const auto &e : c->entries)
        if (e.dev) cudaFree(e.dev);
    delete c;
    m->key_cache = nullptr;
}

extern "C" int pgb_gadget_key_pin(pgb_module *m, const pgb_vmp_pmat *key) {
    PGB_REQUIRE(m && key && key->data, "gadget_key_pin: null argument");
    PGB_REQUIRE(key->n == m->n, "gadget_key_pin: ring degree mismatch");
    if (key_is_pinned(m, key->data)) return PGB_OK;
    cache_of(m, true)->pinned.push_back({(const char *)key->data, pgb_bytes_of_vmp_pmat(m, key->rows, key->cols_in, key->cols_out, key->size)});
    return PGB_OK;
}

extern "C" int pgb_gadget_key_unpin(pgb_module *m, const pgb_vmp_pmat *key) {
    PGB_REQUIRE(m && key && key->data, "gadget_key_unpin: null argument");
    KeyCache *c = m->key_cache;
    if (!c) return PGB_OK;
    for (size_t i = 0; i < c->pinned.size(); i++)
        if (c->pinned[i].first == (const char *)key->data) {
            key_cache_invalidate(m, c->pinned[i].first, c->pinned[i].second);
            c->pinned.erase(c->pinned.begin() + (long)i);
            break;
        }
    return PGB_OK;
}
