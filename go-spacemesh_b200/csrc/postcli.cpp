// postcli.cpp — `b200postcli`: a postcli-compatible initialisation CLI on top of libb200post.so.
//
// The reference deploys POST data with an init container running `postcli` (systest/cluster/nodes.go:990-999):
//   postcli -id <hex pubkey> -commitmentAtxId <hex> -datadir /data -numUnits N -labelsPerUnit L -scryptN 8192
//           -provider 4294967295 -yes
// This tool accepts the same flags (single-dash, Go `flag` style; "-flag value" or "-flag=value") and drives a
// PostSetupManager session (include/b200post_setup.h).  Differences: `-provider` takes a CUDA ordinal or
// "all"; the CPU provider id 4294967295 is refused (the library has no CPU path).
#include <signal.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "../../include/b200post_setup.h"

static volatile int g_cancel = 0;
static void on_signal(int) { g_cancel = 1; }

static bool unhex32(const std::string &s, uint8_t out[32]) {
    if (s.size() != 64) return false;
    for (int i = 0; i < 32; i++) {
        unsigned v;
        if (sscanf(s.c_str() + 2 * i, "%2x", &v) != 1) return false;
        out[i] = (uint8_t)v;
    }
    return true;
}

int main(int argc, char **argv) {
    std::string id, atx, datadir = "./post-data", provider = "0";
    uint64_t num_units = 0, labels_per_unit = 0, scrypt_n = 8192, max_file_size = 4ull << 30, batch = 1ull << 20;
    bool print_providers = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], v;
        while (!a.empty() && a[0] == '-') a.erase(0, 1);
        const size_t eq = a.find('=');
        if (eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
        auto val = [&]() -> std::string { if (eq != std::string::npos) return v; return i + 1 < argc ? argv[++i] : ""; };
        if (a == "id") id = val();
        else if (a == "commitmentAtxId") atx = val();
        else if (a == "datadir") datadir = val();
        else if (a == "numUnits") num_units = strtoull(val().c_str(), nullptr, 10);
        else if (a == "labelsPerUnit") labels_per_unit = strtoull(val().c_str(), nullptr, 10);
        else if (a == "scryptN") scrypt_n = strtoull(val().c_str(), nullptr, 10);
        else if (a == "maxFileSize") max_file_size = strtoull(val().c_str(), nullptr, 10);
        else if (a == "computeBatchSize") batch = strtoull(val().c_str(), nullptr, 10);
        else if (a == "provider") provider = val();
        else if (a == "printProviders") print_providers = true;
        else if (a == "yes") {}
        else { fprintf(stderr, "unknown flag -%s\n", a.c_str()); return 2; }
    }
    if (print_providers) {
        b200post_provider p[16];
        const int n = b200post_providers(p, 16);
        for (int k = 0; k < n && k < 16; k++) printf("{ID: %u, Model: \"%s\", DeviceType: GPU, HBM: %llu}\n", p[k].id, p[k].model, (unsigned long long)p[k].hbm_bytes);
        return 0;
    }
    uint8_t node_id[32], atx_id[32];
    if (!unhex32(id, node_id) || !unhex32(atx, atx_id)) { fprintf(stderr, "-id and -commitmentAtxId must be 32-byte hex strings\n"); return 2; }
    b200post_post_config cfg;
    b200post_default_post_config(&cfg);
    if (labels_per_unit) cfg.labels_per_unit = labels_per_unit;
    cfg.min_num_units = 1; cfg.max_num_units = 1u << 20;
    b200post_setup_opts o;
    b200post_default_setup_opts(&o);
    o.data_dir = datadir.c_str(); o.num_units = (uint32_t)num_units; o.scrypt_n = scrypt_n; o.max_file_size = max_file_size;
    o.compute_batch_size = batch;
    if (provider == "all") o.provider_id = B200POST_PROVIDER_ALL;
    else o.provider_id = (int64_t)strtoull(provider.c_str(), nullptr, 10);
    if (o.provider_id == (int64_t)B200POST_CPU_PROVIDER_ID) { fprintf(stderr, "provider 4294967295 (CPU) is not served: this build has no CPU path\n"); return 2; }

    b200post_setup_manager *mgr = nullptr;
    if (b200post_setup_manager_new(&cfg, &mgr)) { fprintf(stderr, "error: %s\n", b200post_last_error()); return 1; }
    if (int rc = b200post_setup_prepare_initializer(mgr, &o, node_id, atx_id)) { fprintf(stderr, "prepare: %s (%d)\n", b200post_last_error(), rc); return 1; }
    signal(SIGINT, on_signal); signal(SIGTERM, on_signal);
    const uint64_t total = (uint64_t)o.num_units * cfg.labels_per_unit;
    std::thread progress([&] {
        const auto t0 = std::chrono::steady_clock::now();
        b200post_setup_status st;
        uint64_t first = ~0ull;
        do {
            std::this_thread::sleep_for(std::chrono::seconds(2));
            b200post_setup_get_status(mgr, &st);
            if (first == ~0ull) first = st.num_labels_written;
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "\r%llu / %llu labels (%.1f %%), %.0f labels/s   ", (unsigned long long)st.num_labels_written, (unsigned long long)total,
                    100.0 * st.num_labels_written / (double)total, el > 0 ? (st.num_labels_written - first) / el : 0.0);
        } while (st.state == B200POST_SETUP_IN_PROGRESS || st.state == B200POST_SETUP_PREPARED);
        fprintf(stderr, "\n");
    });
    const int rc = b200post_setup_start_session(mgr, &g_cancel);
    progress.join();
    if (rc == B200POST_ERR_CANCELLED) { fprintf(stderr, "stopped; run again to resume\n"); return 130; }
    if (rc) { fprintf(stderr, "init failed: %s (%d)\n", b200post_last_error(), rc); return 1; }
    b200post_post_metadata md;
    if (b200post_load_metadata(datadir.c_str(), &md) == 0 && md.has_nonce) printf("initialization complete; VRF nonce %llu\n", (unsigned long long)md.nonce);
    b200post_setup_manager_free(mgr);
    return 0;
}
