// psychopath_main.cpp — `psychopath` command line: same flags, console output and Blender stream
// protocol as the reference's src/main.rs:62-360, with the render core replaced by the CUDA library
// behind include/psy_cuda.h.  (the reference CLI is Rust, which cannot be built in
// this image, so the CLI is restated in C++ on top of libpsy_host.so + libpsy_cuda.so.)
//
//   -i/--input FILE   -s/--spp N   -b/--spb N   --crop X1 Y1 X2 Y2   -t/--threads N (accepted, unused)
//   --stats  --dev  --serialized_output  --use_stdin        plus  --device N (GPU index)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/psy_cuda.h"

extern "C" {
const char* psyh_last_error(void);
int psyh_count_scenes(const char* text, uint32_t* count);
int psyh_load_psy_string(const char* text, uint32_t scene_index, void** out);
const PsySceneDesc* psyh_scene_desc(void* scene);
const char* psyh_output_path(void* scene);
void psyh_scene_set_render(void* scene, uint32_t res_x, uint32_t res_y, uint32_t spp, uint32_t seed);
void psyh_scene_free(void* scene);
int psyh_write_png(const char* path, const float* xyz, uint32_t w, uint32_t h);
int psyh_write_exr(const char* path, const float* xyz, uint32_t w, uint32_t h);
size_t psyh_bucket_base64(const float* xyz, uint32_t w, uint32_t h, char* out, size_t cap);
void psyh_set_mesh_topology_builder(void* fn, void* ctx, uint32_t min_tris);
void psyh_set_mesh_bvh_builder(void* build, void* take, void* ctx, uint32_t min_tris);
double psyh_take_mesh_topology_seconds(uint32_t* meshes_built);
int psyh_synth_displaced_grid(uint32_t quads_per_side, uint32_t n_vert_keys, uint32_t n_xform_keys, uint32_t res_x,
                              uint32_t res_y, uint32_t spp, uint32_t seed, void** out);
int psyh_synth_instanced(uint32_t asset_quads, uint32_t n_assets, uint32_t n_clusters, uint32_t inst_per_cluster,
                         uint32_t lights_per_cluster, uint32_t cluster_xform_keys, uint32_t res_x, uint32_t res_y,
                         uint32_t spp, uint32_t seed, uint32_t motion_pct, void** out);
}

namespace {

struct Timer {
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    double tick() {
        auto now = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(now - last).count();
        last = now;
        return s;
    }
};

[[noreturn]] void die(const std::string& msg) {
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(1);
}

struct StreamState {
    uint64_t total_pixels = 0, pixels_rendered = 0;
    bool serialized = false;
    std::string last_pct = "0.00%";
};

// renderer.rs:312-346: progress print, or the DIV / % / rect / base64 / BUCKET_END block for Blender
void on_bucket(void* user, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, const float* xyz) {
    StreamState* st = static_cast<StreamState*>(user);
    st->pixels_rendered += (uint64_t)(x1 - x0) * (y1 - y0);
    char pct[32];
    std::snprintf(pct, sizeof(pct), "%.2f%%", (double)st->pixels_rendered / (double)st->total_pixels * 100.0);
    if (st->serialized) {
        size_t need = psyh_bucket_base64(xyz, x1 - x0, y1 - y0, nullptr, 0);
        std::vector<char> buf(need);
        psyh_bucket_base64(xyz, x1 - x0, y1 - y0, buf.data(), buf.size());
        std::printf("DIV\n%s\n%u %u %u %u\n%s\nBUCKET_END\nDIV\n", pct, x0, y0, x1, y1, buf.data());
    } else if (st->last_pct != pct) {
        std::printf("\r%s", pct);
    }
    st->last_pct = pct;
    std::fflush(stdout);
}

}  // namespace

int main(int argc, char** argv) {
    std::string input;
    bool has_input = false, stats = false, dev = false, serialized = false, use_stdin = false, has_crop = false;
    uint32_t spp = 0, spb = 0, crop[4] = {0, 0, 0, 0};
    int device = 0;
    std::vector<int> devices;  // --gpus N / --device-list a,b,c (flags the GPU host adds)
    bool counters = false, gpu_build = false;
    std::string probe_in, probe_out;
    std::string synth;  // --synth NAME: one of BASELINE.json's synthetic configurations instead of a .psy file
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) {
            if (i + n >= argc) die("Argument '" + a + "' needs " + std::to_string(n) + " value(s)");
        };
        if (a == "-i" || a == "--input") {
            need(1);
            input = argv[++i];
            has_input = true;
        } else if (a == "-s" || a == "--spp") {
            need(1);
            spp = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        } else if (a == "-b" || a == "--spb") {
            need(1);
            spb = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        } else if (a == "-t" || a == "--threads") {
            need(1);
            ++i;  // host threads are irrelevant to the GPU render core
        } else if (a == "--crop") {
            need(4);
            for (int k = 0; k < 4; ++k) crop[k] = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
            has_crop = true;
            if (crop[0] > crop[2]) die("Argument '--crop': X1 must be less than or equal to X2");
            if (crop[1] > crop[3]) die("Argument '--crop': Y1 must be less than or equal to Y2");
        } else if (a == "--device") {
            need(1);
            device = std::atoi(argv[++i]);
        } else if (a == "--gpus") {
            need(1);
            int n = std::atoi(argv[++i]);
            if (n < 1) die("Argument '--gpus' must be at least 1");
            devices.clear();
            for (int k = 0; k < n; ++k) devices.push_back(k);
        } else if (a == "--device-list") {
            need(1);
            devices.clear();
            std::stringstream ss(argv[++i]);
            std::string tok;
            while (std::getline(ss, tok, ',')) devices.push_back(std::atoi(tok.c_str()));
            if (devices.empty()) die("Argument '--device-list' needs a comma-separated list of CUDA device indices");
        } else if (a == "--counters") {
            counters = true;
        } else if (a == "--synth") {
            need(1);
            synth = argv[++i];
        } else if (a == "--gpu-build") {
            gpu_build = true;  // mesh BVH order + topology from the GPU LBVH build
        } else if (a == "--probe-rays") {
            need(1);
            probe_in = argv[++i];
        } else if (a == "--probe-out") {
            need(1);
            probe_out = argv[++i];
        } else if (a == "--stats") {
            stats = true;
        } else if (a == "--dev") {
            dev = true;
        } else if (a == "--serialized_output") {
            serialized = true;
        } else if (a == "--use_stdin") {
            use_stdin = true;
        } else {
            die("Unknown argument '" + a + "'");
        }
    }
    if (dev) {
        std::printf("PsyBvh4Node size: %zu bytes\nPsyHit size: %zu bytes\nPsyRay size: %zu bytes\n", sizeof(PsyBvh4Node),
                    sizeof(PsyHit), sizeof(PsyRay));
        return 0;
    }
    if (!has_input && !use_stdin && synth.empty())
        die("error: The following required arguments were not provided: --input <FILE>");
    if (!synth.empty() && synth != "C3" && synth != "C4" && synth != "C5" && synth != "C5r1" && synth != "tiny")
        die("Argument '--synth': one of C3, C4, C5, C5r1, tiny");

    Timer t;
    if (!serialized) std::printf("Parsing scene file...\n");
    std::string text;
    if (!synth.empty()) {
        // (nothing to parse: the scene is generated below)
    } else if (use_stdin) {
        // main.rs:194-226: read until the "__PSY_EOF__" marker
        std::string chunk(4096, '\0');
        for (;;) {
            std::cin.read(&chunk[0], (std::streamsize)chunk.size());
            std::streamsize n = std::cin.gcount();
            if (n <= 0) die("Unexpected end of scene input.");
            size_t from = text.size() < 11 ? 0 : text.size() - 11;
            text.append(chunk.data(), (size_t)n);
            size_t pos = text.find("__PSY_EOF__", from);
            if (pos != std::string::npos) {
                text.resize(pos);
                break;
            }
        }
    } else {
        std::ifstream f(input, std::ios::binary);
        if (!f) die("Cannot open '" + input + "'");
        std::stringstream ss;
        ss << f.rdbuf();
        text = ss.str();
    }
    uint32_t n_scenes = 0;
    if (!synth.empty())
        n_scenes = 1;
    else if (psyh_count_scenes(text.c_str(), &n_scenes) != 0)
        die(std::string("Parse error: ") + psyh_last_error());
    if (!serialized) std::printf("\tParsed scene file in %.3fs\n", t.tick());

    PsyContext* ctx = nullptr;
    if (devices.empty()) devices.push_back(device);
    if (psy_init(devices.data(), (int)devices.size(), &ctx) != PSY_OK) die(std::string("psy_init: ") + psy_last_error());
    if (gpu_build) {
        // whole mesh BVH4s on the device; meshes it refuses (more than 8 time keys) get at least their topology there
        psyh_set_mesh_bvh_builder(reinterpret_cast<void*>(&psy_mesh_bvh_build), reinterpret_cast<void*>(&psy_mesh_bvh_take), ctx,
                                  1u << 16);
        psyh_set_mesh_topology_builder(reinterpret_cast<void*>(&psy_build_mesh_topology), ctx, 1u << 16);
    }
    std::string device_names;
    for (size_t k = 0; k < devices.size(); ++k) device_names += (k ? "," : "") + std::to_string(devices[k]);

    for (uint32_t si = 0; si < n_scenes; ++si) {
        t.tick();
        if (!serialized) std::printf("Building scene...\n");
        void* hs = nullptr;
        if (!synth.empty()) {
            // BASELINE.json's synthetic configurations, generated straight into the flat scene:
            //   C3  10 M-triangle displaced mesh, 4 deformation + 4 transform keys, 1920x1080, 128 spp
            //   C4  100 clusters x (10 instances of 4 x 500k-triangle assets + 10 area lights), 1920x1080, 64 spp
            //   C5  the same with moving clusters (motion-blurred BVH nodes), 3840x2160, 256 spp   (C5r1: round 1's motion)
            //   tiny  a small instanced scene for tests, 320x180, 8 spp
            int rc = 1;
            if (synth == "C3")
                rc = psyh_synth_displaced_grid(2240, 4, 4, 1920, 1080, 128, 3, &hs);
            else if (synth == "C4")
                rc = psyh_synth_instanced(500, 4, 100, 10, 10, 1, 1920, 1080, 64, 4, 100, &hs);
            else if (synth == "C5")
                rc = psyh_synth_instanced(500, 4, 100, 10, 10, 2, 3840, 2160, 256, 5, 1250, &hs);
            else if (synth == "C5r1")
                rc = psyh_synth_instanced(500, 4, 100, 10, 10, 2, 3840, 2160, 256, 5, 100, &hs);
            else if (synth == "tiny")
                rc = psyh_synth_instanced(40, 4, 16, 6, 4, 2, 320, 180, 8, 5, 1250, &hs);
            else
                die("Argument '--synth': one of C3, C4, C5, C5r1, tiny");
            if (rc != 0) die(std::string("Scene generation failed: ") + psyh_last_error());
        } else if (psyh_load_psy_string(text.c_str(), si, &hs) != 0) {
            die(std::string("Parse error: ") + psyh_last_error());
        }
        const PsySceneDesc* desc = psyh_scene_desc(hs);
        if (spp) {
            if (!serialized) std::printf("\tOverriding scene spp: %u\n", spp);
            psyh_scene_set_render(hs, 0, 0, spp, desc->seed);
        }
        PsyScene* scene = nullptr;
        if (psy_scene_create(ctx, desc, &scene) != PSY_OK) die(std::string("psy_scene_create: ") + psy_last_error());
        if (!serialized) std::printf("\tBuilt scene in %.3fs\n", t.tick());
        if (gpu_build && !serialized) {
            uint32_t meshes = 0;
            const double secs = psyh_take_mesh_topology_seconds(&meshes);
            std::printf("\t%u mesh BVHs built on the GPU in %.3fs\n", meshes, secs);
        }
        if (!probe_in.empty()) {
            // --probe-rays FILE: trace a file of PsyRay records (36 bytes each: orig[3] dir[3] time max_t flags) through
            // the scene instead of rendering it, and write the PsyHit records (32 bytes each) to --probe-out (default
            // FILE.hits) — the closest-hit parity check of the north star from the command line
            std::ifstream pf(probe_in, std::ios::binary);
            if (!pf) die("Cannot open '" + probe_in + "'");
            std::vector<char> raw((std::istreambuf_iterator<char>(pf)), std::istreambuf_iterator<char>());
            if (raw.size() % sizeof(PsyRay) != 0) die("'" + probe_in + "' is not a whole number of 36-byte ray records");
            const size_t n = raw.size() / sizeof(PsyRay);
            std::vector<PsyHit> hits(n);
            if (psy_trace_rays(scene, reinterpret_cast<const PsyRay*>(raw.data()), n, hits.data(), nullptr) != PSY_OK)
                die(std::string("psy_trace_rays: ") + psy_last_error());
            const std::string outp = probe_out.empty() ? probe_in + ".hits" : probe_out;
            std::ofstream of(outp, std::ios::binary);
            of.write(reinterpret_cast<const char*>(hits.data()), (std::streamsize)(n * sizeof(PsyHit)));
            if (!of) die("Cannot write '" + outp + "'");
            size_t n_hit = 0;
            for (auto& h : hits) n_hit += (h.kind != PSY_HIT_MISS && h.kind != PSY_HIT_OCCLUDED) ? 1 : 0;
            std::printf("Traced %zu probe rays (%zu hits) in %.3fs, wrote '%s'\n", n, n_hit, t.tick(), outp.c_str());
            psy_scene_destroy(scene);
            psyh_scene_free(hs);
            continue;
        }
        if (!serialized) std::printf("Rendering scene on CUDA device%s %s...\n", devices.size() > 1 ? "s" : "", device_names.c_str());

        PsyRenderParams params;
        std::memset(&params, 0, sizeof(params));
        params.crop_enabled = has_crop ? 1u : 0u;
        params.crop_x1 = crop[0], params.crop_y1 = crop[1], params.crop_x2 = crop[2], params.crop_y2 = crop[3];
        // --spb on the reference bounds the paths in flight per thread (default 4096); on the GPU the same
        // knob sizes the device path pool, whose default is much larger.
        // (without --spb the library chooses: the whole image up to 64 Mi paths on one GPU, 12 buckets per GPU when the
        // image is shared, 4 Mi-path buckets when Blender displays them as they arrive)
        params.max_samples_per_bucket = spb;  // 0 = the library's default for the number of GPUs / streaming
        params.collect_counters = (stats || counters) ? 1u : 0u;
        std::vector<float> image((size_t)desc->res_x * desc->res_y * 3);
        PsyStats rs;
        StreamState st;
        st.serialized = serialized;
        uint32_t cw = desc->res_x, chh = desc->res_y;
        if (has_crop) {
            uint32_t x1 = std::min(crop[0], desc->res_x - 1), y1 = std::min(crop[1], desc->res_y - 1);
            uint32_t x2 = std::min(crop[2], desc->res_x - 1), y2 = std::min(crop[3], desc->res_y - 1);
            cw = x2 - x1 + 1, chh = y2 - y1 + 1;
        }
        st.total_pixels = (uint64_t)cw * chh;
        std::printf("0.00%%");
        std::fflush(stdout);
        if (psy_render_image(scene, &params, image.data(), &rs, on_bucket, &st) != PSY_OK)
            die(std::string("psy_render_image: ") + psy_last_error());
        std::printf("\r                \r");

        if (!serialized) {
            double rtime = t.tick();
            std::printf("\tRendered scene in %.3fs\n", rtime);
            std::printf("\t\tTrace:                  %.3fs\n", rs.trace_time);
            std::printf("\t\t\tRays traced:          %llu\n", (unsigned long long)rs.ray_count);
            std::printf("\t\t\tRays/sec:             %llu\n",
                        (unsigned long long)(rs.trace_time > 0 ? (double)rs.ray_count / rs.trace_time : 0.0));
            std::printf("\t\t\tRay/node tests:       %llu\n", (unsigned long long)rs.accel_node_visits);
            std::printf("\t\tInitial ray generation: %.3fs\n", rs.initial_ray_generation_time);
            std::printf("\t\tRay generation:         %.3fs\n", rs.ray_generation_time);
            std::printf("\t\tSample writing:         %.3fs\n", rs.sample_writing_time);
            if (counters) {
                // --counters: what the roofline is made of, per ray
                const double rays = (double)std::max<uint64_t>(rs.ray_count, 1);
                const double nn = rs.accel_node_visits / rays, nm = rs.motion_node_visits / rays, nt = rs.tri_tests / rays,
                             ni = rs.instance_visits / rays;
                std::printf("\t\tPer ray: %.2f node visits (%.2f with two time keys), %.2f triangle tests, %.2f instance entries\n",
                            nn, nm, nt, ni);
                std::printf("\t\tBuckets: %llu, kernel launches: %llu\n", (unsigned long long)rs.buckets,
                            (unsigned long long)rs.kernel_launches);
            }

            std::string out_path = psyh_output_path(hs);
            std::printf("Writing image to disk into '%s'...\n", out_path.c_str());
            int rc;
            if (out_path.size() >= 4 && out_path.compare(out_path.size() - 4, 4, ".png") == 0)
                rc = psyh_write_png(out_path.c_str(), image.data(), desc->res_x, desc->res_y);
            else if (out_path.size() >= 4 && out_path.compare(out_path.size() - 4, 4, ".exr") == 0)
                rc = psyh_write_exr(out_path.c_str(), image.data(), desc->res_x, desc->res_y);
            else
                die("Unknown output file extension.");
            if (rc != 0) die("Failed to write image...");
            std::printf("\tWrote image in %.3fs\n", t.tick());
        }
        psy_scene_destroy(scene);
        psyh_scene_free(hs);
    }
    psy_shutdown(ctx);
    std::printf("\n");
    return 0;
}
