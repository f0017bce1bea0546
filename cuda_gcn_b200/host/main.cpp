// main.cpp — the `./gcn-cuda <dataset>` CLI (reference src/main.cpp:15-49): same usage line, same
// "Cannot read input" failure, same stdout lines.  The nine optional positionals the reference's usage
// text advertises but never parses (main.cpp:24-25) ARE parsed here: hidden_dim dropout learning_rate
// weight_decay epochs early_stopping take effect; num_nodes/input_dim/output_dim stay parser-derived
// ("-" keeps a default).  `gcn-cuda synth:<preset>[:scale]` runs a generated dataset instead of files.
// Environment: GCN_SEED, GCN_PLAN=auto|modules|fused, GCN_DATA_DIR, GCN_PROFILE=1, GCN_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "check.h"
#include "gcn.h"
#include "parser.h"
#include "synth.h"
#include "timer.h"

int main(int argc, char **argv) {
    setbuf(stdout, NULL);
    if (argc < 2) {
        std::cout << "gcn-cuda graph_name [num_nodes input_dim hidden_dim "
                     "output_dim dropout learning_rate, weight_decay epochs early_stopping]"
                  << std::endl;
        return EXIT_FAILURE;
    }
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
    GCNK_CHECK(gcnk_set_device(dev && *dev ? atoi(dev) : 0));
    const char *prof = getenv("GCN_PROFILE");
    const bool profile = prof && *prof && strcmp(prof, "0");
    gpu_timer_enable(profile);

    std::cout << "RUNNING ON GPU" << std::endl;
    GCN gcn(params, &data);
    gcn.run();
    if (profile)
        for (int t = TMR_MATMUL_FW; t < __NUM_TMR; t++)
            if (timer_calls((timer_instance)t))
                printf("%-12s calls=%d total=%.3fms avg=%.3fus\n", timer_name((timer_instance)t), timer_calls((timer_instance)t),
                       timer_total((timer_instance)t) * 1e3, timer_total((timer_instance)t) * 1e6 / timer_calls((timer_instance)t));
    return EXIT_SUCCESS;
}
