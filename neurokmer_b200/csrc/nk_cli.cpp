// nk_cli.cpp — `neurokmer` command line over the C ABI.
//
// Flag surface and result block of the reference binary (src/main.rs:10-25, :49-74):
//   neurokmer --input/-i FILE [--k/-k 31] [--pool-size 1000000] [--canonical] [--streaming]
// Additive flags with the reference's constants as defaults: --steps 1000 (spiking_hash.rs:70),
// --top-n 20 (main.rs:50), --device 0, --gpus N (shard the input over N GPUs of this process: nk_create_multi).  The "unique k-mers colliding" column (kmer_per_neuron, main.rs:54-61)
// is computed for the printed rows by a second read of the file (nk_set_file_uniques); --exact builds the
// whole exact k-mer side table instead (O(windows) memory), --no-uniques skips the column ("n/a", never a guess).
// LIF constants are the ones main.rs:37 hard-codes (threshold 1.0, leak 0.95, refractory 2, cost 1.0).
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/neurokmer.h"

static void usage() {
    fprintf(stderr,
            "Neuromorphic k-mer counting with fixed-size spiking neuron pool (B200 / sm_100a)\n\n"
            "Usage: neurokmer --input <INPUT> [OPTIONS]\n\n"
            "  -i, --input <INPUT>          FASTA/FASTQ file\n"
            "  -k, --k <K>                  k-mer length, 1..32 [default: 31]\n"
            "      --pool-size <POOL_SIZE>  neurons in the pool [default: 1000000]\n"
            "      --canonical              count min(forward, reverse complement)\n"
            "      --streaming              streaming driver (process_file_streaming)\n"
            "      --steps <N>              LIF ticks [default: 1000]\n"
            "      --top-n <N>              rows of the result block [default: 20]\n"
            "      --device <ID>            CUDA device ordinal [default: 0]\n"
            "      --gpus <N>               shard the input over GPUs 0..N-1 (0 = all) [default: 1]\n"
            "      --devices <a,b,..>       shard the input over exactly these CUDA devices\n"
            "      --exact                  build the exact k-mer table (uniques column from it; with several GPUs the\n"
            "                               per-GPU tables are merged by neuron slice)\n"
            "      --no-uniques             skip the second read of the file that fills the uniques column\n");
}

// Rust's `{}` for f64 prints the shortest representation that round-trips, without a
// trailing ".0" only when... it DOES print "76082638" for 76082638.0 (Display for floats
// prints integers without a fraction).  Reproduce: integral values print as integers.
static std::string rust_f64(double v) {
    char buf[64];
    if (v == (double)(long long)v && v > -1e15 && v < 1e15) {
        snprintf(buf, sizeof buf, "%lld", (long long)v);
        return buf;
    }
    for (int prec = 1; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    return buf;
}

int main(int argc, char** argv) {
    std::string input;
    nk_config cfg;
    nk_config_default(&cfg);
    uint64_t top_n = 20;
    int streaming = 0, exact = 0, timing = 0, no_uniques = 0, gpus = 1;
    std::vector<int32_t> devices;
    const auto T0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count(); };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") input = val("--input");
        else if (a == "-k" || a == "--k") cfg.k = (uint32_t)strtoul(val("--k"), nullptr, 10);
        else if (a == "--pool-size") cfg.pool_size = strtoull(val("--pool-size"), nullptr, 10);
        else if (a == "--canonical") cfg.use_canonical = 1;
        else if (a == "--streaming") streaming = 1;
        else if (a == "--steps") cfg.steps = strtoull(val("--steps"), nullptr, 10);
        else if (a == "--top-n") top_n = strtoull(val("--top-n"), nullptr, 10);
        else if (a == "--device") cfg.device = atoi(val("--device"));
        else if (a == "--gpus") gpus = atoi(val("--gpus"));
        else if (a == "--devices") {
            for (const char* p = val("--devices"); *p;) {
                devices.push_back((int32_t)strtol(p, const_cast<char**>(&p), 10));
                if (*p == ',') ++p; else if (*p) { fprintf(stderr, "error: --devices takes a comma-separated list of ordinals\n"); return 2; }
            }
        }
        else if (a == "--exact") exact = 1;
        else if (a == "--no-uniques") no_uniques = 1;
        else if (a == "--timing") timing = 1;  // phase wall times on stderr
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "error: unexpected argument '%s'\n\n", a.c_str()); usage(); return 2; }
    }
    if (input.empty()) { fprintf(stderr, "error: the following required arguments were not provided:\n  --input <INPUT>\n\n"); usage(); return 2; }

    nk_counter* h = nullptr;
    if (gpus == 0) {
        int32_t nd = 0;
        nk_device_count(&nd);
        gpus = nd > 0 ? nd : 1;
    }
    if (gpus < 0) { fprintf(stderr, "error: --gpus must be >= 0\n"); return 2; }
    if (!devices.empty()) gpus = (int)devices.size();
    const int crc = !devices.empty() ? nk_create_multi(&cfg, devices.data(), (int32_t)devices.size(), &h)
                    : gpus > 1       ? nk_create_multi(&cfg, nullptr, gpus, &h)
                                     : nk_create(&cfg, &h);
    if (crc != NK_OK) { fprintf(stderr, "Error: %s\n", nk_last_error()); return 1; }
    if (timing) fprintf(stderr, "[timing] nk_create done at %.3f s\n", since());
    if (exact && nk_enable_exact_counts(h, 1) != NK_OK) { fprintf(stderr, "Error: %s\n", nk_last_error()); return 1; }
    if (!exact && !no_uniques && top_n >= 1 && top_n <= 2048 && nk_set_file_uniques(h, top_n) != NK_OK) {
        fprintf(stderr, "Error: %s\n", nk_last_error());
        return 1;
    }
    if (nk_process_file(h, input.c_str(), streaming) != NK_OK) {
        fprintf(stderr, "Error: %s\n", nk_last_error());  // the reference returns Err from main
        nk_destroy(h);
        return 1;
    }

    if (timing) {
        nk_timings t;
        nk_last_timings(h, &t);
        fprintf(stderr, "[timing] nk_process_file done at %.3f s (device: h2d %.2f ms, mark %.2f, count %.2f, post %.2f; %llu k-mers)\n",
                since(), t.h2d_ms, t.mark_ms, t.count_ms, t.lif_ms, (unsigned long long)t.kmers);
    }
    printf("\n=== Top 20 Abundant Neuron Groups (Highest Spike Rates) ===\n");
    std::vector<nk_top_entry> top(top_n ? top_n : 1);
    uint64_t got = 0;
    if (nk_top_n(h, top_n, top.data(), &got) != NK_OK) { fprintf(stderr, "Error: %s\n", nk_last_error()); return 1; }
    if (got == 0) {
        printf("No spikes fired (empty file or too small k)\n");
    } else {
        for (uint64_t r = 0; r < got; ++r) {
            // "{:3}: Neuron {:6} → {:8} spikes ({} unique k-mers colliding)"
            if (top[r].uniques == NK_UNIQUES_NOT_COMPUTED)
                printf("%3" PRIu64 ": Neuron %6" PRIu64 " \xE2\x86\x92 %8" PRIu64 " spikes (n/a unique k-mers colliding)\n", r + 1,
                       top[r].idx, top[r].spikes);
            else
                printf("%3" PRIu64 ": Neuron %6" PRIu64 " \xE2\x86\x92 %8" PRIu64 " spikes (%u unique k-mers colliding)\n", r + 1,
                       top[r].idx, top[r].spikes, top[r].uniques);
        }
    }
    uint64_t spikes = 0;
    double energy = 0.0;
    nk_total_spikes(h, &spikes);
    nk_energy_used(h, &energy);
    printf("\nTotal spikes fired: %" PRIu64 "\n", spikes);
    printf("Simulated energy used: %s\n", rust_f64(energy).c_str());
    printf("Neuron pool size used: %" PRIu64 "\n", (uint64_t)cfg.pool_size);
    printf("Streaming mode: %s\n", streaming ? "true" : "false");
    nk_destroy(h);
    if (timing) fprintf(stderr, "[timing] exit at %.3f s\n", since());
    return 0;
}
