// prv_simulation.cpp -- command-line driver with the reference's stdin protocol (main.cpp:2294-2309: a mode integer,
// then object names until "-1") for the mode the hot path serves: GetCoverage (3), the dataset-generation loop of
// main.cpp:2343-2462 -- for every object, view sets N = 3, 3+add, ... <= coverage_view_num_max and then N = 100, each
// skipped when <gt_path>/<N>.json already exists (idempotent resume), images + transforms written per view set.
//
//   prv_simulation [config.yaml] [--device D] [--no-coverage] [--rank R --world G]     (objects are sharded o mod G == R)
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "NBV_Net_Labeler.hpp"

#define ViewCover 0
#define ViewNovel 1
#define GetSizeTest 2
#define GetCoverage 3
#define InstantNGP 4

static bool view_set_done(const std::shared_ptr<Share_Data>& sd, int n) {
    std::ifstream fin_json(sd->gt_path + "/" + std::to_string(n) + ".json");
    return fin_json.is_open();
}

int main(int argc, char** argv) {
    std::string config = "../DefaultConfiguration.yaml";
    int device = 0, rank = 0, world = 1;
    bool coverage = true;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--rank") && i + 1 < argc) rank = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--world") && i + 1 < argc) world = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--no-coverage")) coverage = false;
        else config = argv[i];
    }
    // (the reference seeds with clock(), Share_Data.hpp:514; PRV_SIM_SEED pins the size draw of main.cpp:866-870 for tests)
    srand(std::getenv("PRV_SIM_SEED") ? (unsigned)std::atoi(std::getenv("PRV_SIM_SEED")) : (unsigned)time(0));
    int mode;
    std::cout << "input mode:";
    if (!(std::cin >> mode)) return 1;
    std::vector<std::string> names;
    std::cout << "input models:" << std::endl;
    std::string name;
    while (std::cin >> name) {
        if (name == "-1") break;
        names.push_back(name);
    }
    if (mode != GetCoverage) {
        std::cout << "mode " << mode << " is outside the ray-cast / coverage / render hot path served by this build (only mode 3, GetCoverage)." << std::endl;
        return 2;
    }
    int failures = 0;
    for (size_t i = 0; i < names.size(); i++) {
        if ((int)(i % (size_t)world) != rank) continue;  // objects shard across GPUs, no collective
        std::shared_ptr<Share_Data> share_data = std::make_shared<Share_Data>(config, names[i], -1);
        if (!share_data->config_loaded) return 1;
        std::shared_ptr<NBV_Net_Labeler> labeler = std::make_shared<NBV_Net_Labeler>(share_data, 0, 0, device);
        if (!labeler->object_is_ok_size) {
            failures++;
            continue;
        }
        const clock_t t0 = clock();
        std::vector<int> sets;
        for (int n = 3; n <= share_data->coverage_view_num_max; n += share_data->coverage_view_num_add) sets.push_back(n);
        sets.push_back(100);
        for (int n : sets) {
            if (view_set_done(share_data, n)) continue;
            labeler->set_view_count(n);
            if ((int)labeler->view_space->views.size() != n) {
                std::cout << "view space " << n << " not available. Skip." << std::endl;
                continue;
            }
            if (labeler->get_coverage(coverage) != 0) failures++;
        }
        std::cout << "images get with executed time " << (double)(clock() - t0) / CLOCKS_PER_SEC * 1000.0 << " ms." << std::endl;
    }
    return failures ? 3 : 0;
}
