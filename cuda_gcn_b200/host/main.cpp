// main.cpp — the `./gcn-cuda <dataset>` CLI (reference src/main.cpp:15-49): same usage line, same
// "Cannot read input" failure, same stdout lines.  The nine optional positionals the reference's usage
// text advertises but never parses (main.cpp:24-25) ARE parsed here: hidden_dim dropout learning_rate
// weight_decay epochs early_stopping take effect; num_nodes/input_dim/output_dim stay parser-derived
// ("-" keeps a default).  `gcn-cuda synth:<preset>[:scale]` runs a generated dataset instead of files.
// Environment: GCN_SEED, GCN_PLAN=auto|modules|fused, GCN_DATA_DIR, GCN_PROFILE=1, GCN_DEVICE.
// GCN_GPUS=N (N = 2..8) trains row-partitioned on N GPUs of this node: the process forks one worker per GPU before any
// CUDA call (one process per GPU, as NCCL and CUDA IPC want), rank 0 hands the 128-byte NCCL id to the others over
// pipes created before the fork, every worker loads the dataset (the .gcnbin cache makes that cheap) and builds the
// same engine with its rank; rank 0 prints the reference's lines.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include "check.h"
#include "gcn.h"
#include "parser.h"
#include "synth.h"
#include "timer.h"

// one worker per GPU: returns this process's rank (0 in the parent when world == 1); fills id128 on every rank
static int fork_workers(int world, unsigned char *id128, std::vector<pid_t> *children) {
    // pipes[r]: rank 0 -> rank r
    std::vector<int> rd(world, -1), wr(world, -1);
    for (int r = 1; r < world; r++) {
        int fd[2];
        if (pipe(fd) != 0) { perror("pipe"); exit(EXIT_FAILURE); }
        rd[r] = fd[0]; wr[r] = fd[1];
    }
    int rank = 0;
    for (int r = 1; r < world; r++) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); exit(EXIT_FAILURE); }
        if (pid == 0) { rank = r; children->clear(); break; }
        children->push_back(pid);
    }
    if (rank == 0) {
        for (int r = 1; r < world; r++) close(rd[r]);
        GCNK_CHECK(gcnk_comm_unique_id(id128));                 // no CUDA context needed, and none exists before the fork
        for (int r = 1; r < world; r++) {
            if (write(wr[r], id128, 128) != 128) { perror("write"); exit(EXIT_FAILURE); }
            close(wr[r]);
        }
    } else {
        for (int r = 1; r < world; r++) { close(wr[r]); if (r != rank) close(rd[r]); }
        size_t got = 0;
        while (got < 128) {
            const ssize_t k = read(rd[rank], id128 + got, 128 - got);
            if (k <= 0) { fprintf(stderr, "gcn-cuda: rank %d did not receive the NCCL id\n", rank); exit(EXIT_FAILURE); }
            got += (size_t)k;
        }
        close(rd[rank]);
    }
    return rank;
}

int main(int argc, char **argv) {
    setbuf(stdout, NULL);
    if (argc < 2) {
        std::cout << "gcn-cuda graph_name [num_nodes input_dim hidden_dim "
                     "output_dim dropout learning_rate, weight_decay epochs early_stopping]"
                  << std::endl;
        return EXIT_FAILURE;
    }
    const char *gp = getenv("GCN_GPUS");
    const int world = gp && *gp ? atoi(gp) : 1;
    if (world < 1 || world > 8) { std::cerr << "GCN_GPUS must be 1..8" << std::endl; return EXIT_FAILURE; }
    unsigned char id128[128] = {0};
    std::vector<pid_t> children;
    int rank = 0;
    if (world > 1) {
        if (!getenv("GCN_SEED")) {                              // every rank must draw the same weights and masks
            char buf[32];
            snprintf(buf, sizeof buf, "%ld", (long)time(NULL));
            setenv("GCN_SEED", buf, 1);
        }
        rank = fork_workers(world, id128, &children);
    }
    const bool chatty = rank == 0;

    GCNParams params = GCNParams::get_default();
    GCNData data;
    const std::string input_name(argv[1]);
    if (input_name.rfind("synth:", 0) == 0) {
        std::string preset = input_name.substr(6);
        double scale = 1.0;
        const size_t colon = preset.find(':');
        if (colon != std::string::npos) { scale = atof(preset.c_str() + colon + 1); preset.resize(colon); }
        SynthSpec spec;
        if (!synth_preset(preset.c_str(), scale, &spec) || !synth_generate(spec, &params, &data)) {
            std::cerr << "Cannot read input: " << input_name << std::endl;
            exit(EXIT_FAILURE);
        }
        if (preset == "products") params.hidden_dim = 256;
    } else {
        Parser parser(&params, &data, input_name);
        parser.set_quiet(!chatty);
        parser.set_cache_write(chatty);                         // rank 0 alone writes data/<name>.gcnbin
        if (!parser.parse()) {
            std::cerr << "Cannot read input: " << input_name << std::endl;
            exit(EXIT_FAILURE);
        }
    }
    auto arg = [&](int i) -> const char * { return argc > i && strcmp(argv[i], "-") ? argv[i] : nullptr; };
    if (arg(4)) params.hidden_dim = atoi(arg(4));
    if (arg(6)) params.dropout = (float)atof(arg(6));
    if (arg(7)) params.learning_rate = (float)atof(arg(7));
    if (arg(8)) params.weight_decay = (float)atof(arg(8));
    if (arg(9)) params.epochs = atoi(arg(9));
    if (arg(10)) params.early_stopping = atoi(arg(10));

    const char *dev = getenv("GCN_DEVICE");
    const int device = (dev && *dev ? atoi(dev) : 0) + rank;
    GCNK_CHECK(gcnk_set_device(device));
    const char *prof = getenv("GCN_PROFILE");
    const bool profile = prof && *prof && strcmp(prof, "0");
    gpu_timer_enable(profile);

    if (chatty) std::cout << "RUNNING ON GPU" << std::endl;
    int status = EXIT_SUCCESS;
    if (world > 1) {
        GCNDist dist;
        dist.rank = rank; dist.world = world;
        GCNK_CHECK(gcnk_comm_create(&dist.comm, id128, rank, world, device));
        {
            GCN gcn(params, &data, PLAN_FUSED, !chatty, dist);
            gcn.run();
        }
        GCNK_CHECK(gcnk_comm_destroy(dist.comm));
        for (pid_t pid : children) {                            // rank 0 reaps the workers
            int st = 0;
            if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) status = EXIT_FAILURE;
        }
        if (!chatty) _exit(EXIT_SUCCESS);
        return status;
    }
    GCN gcn(params, &data);
    gcn.run();
    if (profile)
        for (int t = TMR_MATMUL_FW; t < __NUM_TMR; t++)
            if (timer_calls((timer_instance)t))
                printf("%-12s calls=%d total=%.3fms avg=%.3fus\n", timer_name((timer_instance)t), timer_calls((timer_instance)t),
                       timer_total((timer_instance)t) * 1e3, timer_total((timer_instance)t) * 1e6 / timer_calls((timer_instance)t));
    return EXIT_SUCCESS;
}
