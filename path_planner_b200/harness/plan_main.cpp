// ppe_plan_main -- standalone, ROS-free driver of the planner on the B200 edge engine: scenario file in, plan
// (Plan.msg text) + Planner::Stats (one JSON line) out.  It stands where path_planner_node + Executive stand in the
// reference (path_planner_node.cpp:68-112, executive.cpp:79-279): builds the world, runs `cycles` planning cycles of
// Planner::plan, feeding each plan back as previousPlan from the state one `period` later along it.
//
//   ppe_plan_main scenario.txt [--device D] [--plan-out plan.yaml] [--visualize dump.txt]
//
// Scenario file: one directive per line, '#' comments.
//   config <name> <value>        max_speed slow_speed turning_radius coverage_turning_radius time_horizon time_minimum
//                                collision_checking_increment ribbon_width branching_factor heuristic(0..4)
//   map none | map gridworld <file>          GridWorldMap text format (GridWorldMap.cpp:10-82)
//   ribbon <x1> <y1> <x2> <y2>
//   obstacle_binary <x> <y> <heading> <speed> <time> <width> <length>
//   obstacle_gaussian <x> <y> <heading> <speed> <time>
//   start <x> <y> <heading> <speed> <time>
//   budget <seconds>             per cycle (Executive: 0.85, executive.h:183)
//   cycles <n>   period <seconds>   initial_samples <n>   brown_paths <0|1>   frontier <m>
//   virtual_clock <clock0> <tick> [sample_tick]   deterministic now() for reproducible runs
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ppe_harness.h"

static int die(const std::string& msg) {
    std::cerr << "ppe_plan_main: " << msg << std::endl;
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 2) return die("usage: ppe_plan_main scenario.txt [--device D] [--plan-out file] [--visualize file]");
    int device = 0;
    std::string planOut, visOut;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--plan-out" && i + 1 < argc) planOut = argv[++i];
        else if (a == "--visualize" && i + 1 < argc) visOut = argv[++i];
        else return die("unknown argument " + a);
    }
    std::ifstream in(argv[1]);
    if (!in) return die(std::string("cannot read ") + argv[1]);

    ppe_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.max_speed = 2.5; cfg.slow_speed = 0.5; cfg.turning_radius = 8; cfg.coverage_turning_radius = 16; // PlannerConfig.h:179-189
    cfg.time_horizon = 30; cfg.time_minimum = 5; cfg.collision_checking_increment = 0.05;
    cfg.ribbon_width = 2; cfg.collision_penalty_factor = 600; cfg.time_penalty_factor = 1;
    cfg.heuristic = PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K; // Executive's default (executive.cpp:391)
    cfg.branching_factor = 9;
    cfg.tsp_k = 2;
    std::string mapKind = "none", mapFile;
    std::vector<double> ribbons, bin[7], gau[5];
    double start[5] = {0, 0, 0, 2.5, 1};
    pph_plan_options opt;
    std::memset(&opt, 0, sizeof opt);
    opt.time_remaining = 0.85; opt.initial_samples = 100; opt.frontier = -1;
    int cycles = 1;
    double period = 1.0;

    std::string line;
    int lineNo = 0;
    while (std::getline(in, line)) {
        lineNo++;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        std::istringstream ls(line);
        std::string key;
        if (!(ls >> key)) continue;
        auto bad = [&]() { return die("line " + std::to_string(lineNo) + ": malformed '" + key + "'"); };
        if (key == "config") {
            std::string name; double v;
            if (!(ls >> name >> v)) return bad();
            if (name == "max_speed") cfg.max_speed = v; else if (name == "slow_speed") cfg.slow_speed = v;
            else if (name == "turning_radius") cfg.turning_radius = v; else if (name == "coverage_turning_radius") cfg.coverage_turning_radius = v;
            else if (name == "time_horizon") cfg.time_horizon = v; else if (name == "time_minimum") cfg.time_minimum = v;
            else if (name == "collision_checking_increment") cfg.collision_checking_increment = v;
            else if (name == "ribbon_width") cfg.ribbon_width = v; else if (name == "branching_factor") cfg.branching_factor = (int)v;
            else if (name == "heuristic") cfg.heuristic = (int)v; else return bad();
        } else if (key == "map") {
            if (!(ls >> mapKind)) return bad();
            if (mapKind == "gridworld" && !(ls >> mapFile)) return bad();
        } else if (key == "ribbon") {
            double v[4];
            if (!(ls >> v[0] >> v[1] >> v[2] >> v[3])) return bad();
            ribbons.insert(ribbons.end(), v, v + 4);
        } else if (key == "obstacle_binary") {
            double v[7];
            for (double& q : v) if (!(ls >> q)) return bad();
            for (int k = 0; k < 7; k++) bin[k].push_back(v[k]);
        } else if (key == "obstacle_gaussian") {
            double v[5];
            for (double& q : v) if (!(ls >> q)) return bad();
            for (int k = 0; k < 5; k++) gau[k].push_back(v[k]);
        } else if (key == "start") {
            for (double& q : start) if (!(ls >> q)) return bad();
        } else if (key == "budget") { if (!(ls >> opt.time_remaining)) return bad();
        } else if (key == "cycles") { if (!(ls >> cycles)) return bad();
        } else if (key == "period") { if (!(ls >> period)) return bad();
        } else if (key == "initial_samples") { if (!(ls >> opt.initial_samples)) return bad();
        } else if (key == "brown_paths") { if (!(ls >> opt.use_brown_paths)) return bad();
        } else if (key == "frontier") { if (!(ls >> opt.frontier)) return bad();
        } else if (key == "virtual_clock") { if (!(ls >> opt.clock0 >> opt.tick)) return bad(); ls >> opt.sample_tick;
        } else return bad();
    }

    pph_ctx* ctx = nullptr;
    if (pph_create(device, &ctx) != PPE_OK) return die("no usable sm_100 CUDA device: the engine has no CPU path");
    auto ck = [&](int rc, const char* what) { if (rc < 0) { die(std::string(what) + ": " + pph_last_error(ctx)); exit(3); } };
    ck(pph_set_config(ctx, &cfg), "pph_set_config");
    if (mapKind == "gridworld") ck(pph_load_gridworld_map(ctx, mapFile.c_str()), "pph_load_gridworld_map");
    else ck(pph_set_map_none(ctx), "pph_set_map_none");
    if (!bin[0].empty())
        ck(pph_set_obstacles_binary(ctx, (int)bin[0].size(), bin[0].data(), bin[1].data(), bin[2].data(), bin[3].data(), bin[4].data(),
                                    bin[5].data(), bin[6].data()), "pph_set_obstacles_binary");
    else if (!gau[0].empty())
        ck(pph_set_obstacles_gaussian(ctx, (int)gau[0].size(), gau[0].data(), gau[1].data(), gau[2].data(), gau[3].data(), gau[4].data(),
                                      nullptr), "pph_set_obstacles_gaussian");
    else ck(pph_set_obstacles_none(ctx), "pph_set_obstacles_none");
    ck(pph_set_ribbons(ctx, (int)(ribbons.size() / 4), ribbons.data()), "pph_set_ribbons");
    if (!visOut.empty()) { opt.visualize = 1; opt.visualization_path = visOut.c_str(); }

    std::vector<pph_dubins_path> plan(256), previous;
    for (int c = 0; c < cycles; c++) {
        pph_stats st;
        const int n = pph_plan(ctx, start, previous.data(), (int)previous.size(), &opt, plan.data(), (int)plan.size(), &st);
        ck(n, "pph_plan");
        printf("{\"cycle\": %d, \"paths\": %d, \"plan_f\": %.17g, \"plan_h\": %.17g, \"plan_time_penalty\": %.17g, "
               "\"plan_collision_penalty\": %.17g, \"plan_depth\": %llu, \"plan_endtime\": %.17g, \"samples\": %llu, \"generated\": %llu, "
               "\"expanded\": %llu, \"iterations\": %llu, \"now_calls\": %llu, \"true_cost_edges\": %llu, \"dubins_solves\": %llu, "
               "\"engine_batches\": %llu, \"frontier_vertices\": %llu, \"frontier_hits\": %llu, \"exact_expansions\": %llu, "
               "\"wall_seconds\": %.6f, \"seconds_engine_expand\": %.6f, \"seconds_replay\": %.6f, \"seconds_add_samples\": %.6f, "
               "\"seconds_exact\": %.6f}\n",
               c, n, st.plan_f, st.plan_h, st.plan_time_penalty, st.plan_collision_penalty, (unsigned long long)st.plan_depth,
               st.plan_endtime, (unsigned long long)st.samples, (unsigned long long)st.generated, (unsigned long long)st.expanded,
               (unsigned long long)st.iterations, (unsigned long long)st.now_calls, (unsigned long long)st.true_cost_edges,
               (unsigned long long)st.dubins_solves, (unsigned long long)st.engine_batches, (unsigned long long)st.frontier_vertices,
               (unsigned long long)st.frontier_hits, (unsigned long long)st.exact_expansions, st.wall_seconds, st.seconds_engine_expand,
               st.seconds_replay, st.seconds_add_samples, st.seconds_exact);
        if (n == 0) break;
        previous.assign(plan.begin(), plan.begin() + (n < (int)plan.size() ? n : (int)plan.size()));
        if (!planOut.empty()) ck(pph_write_plan_msg(previous.data(), (int)previous.size(), planOut.c_str()), "pph_write_plan_msg");
        if (c + 1 < cycles) {
            double next[5];
            if (pph_advance(ctx, start[4] + period, next) != PPE_OK) break; // the plan ends before the next cycle
            std::memcpy(start, next, sizeof start);
            opt.clock0 += 7; // another seed for the next cycle's sampler when the clock is virtual
        }
    }
    pph_destroy(ctx);
    return 0;
}
