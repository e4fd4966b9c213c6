// example_main.cpp — the reference's start-up + frame loop (P/tutorials.cpp:27-42, P/simpleguidx11.cpp:223-334) reduced
// to the ReSTIR path, on the C++ mirror: build a small room with one ceiling light, render a few frames with temporal
// and spatial reuse, print the image mean the way the reference's statistics panel does (P/simpleguidx11.cpp:308-329).
#include <cstdio>
#include <vector>

#include "restir_b200.hpp"

using namespace rb200;

static void quad(Surface& s, vec3 o, vec3 du, vec3 dv, vec3 n) {
  const vec3 p00 = o, p10{o.x + du.x, o.y + du.y, o.z + du.z}, p01{o.x + dv.x, o.y + dv.y, o.z + dv.z},
             p11{p10.x + dv.x, p10.y + dv.y, p10.z + dv.z};
  const vec3 tri[6] = {p00, p10, p11, p00, p11, p01};
  for (const vec3& p : tri) {
    s.positions.insert(s.positions.end(), {p.x, p.y, p.z});
    s.normals.insert(s.normals.end(), {n.x, n.y, n.z});
  }
}

int main() {
  try {
    const int W = 320, H = 180;
    Scene scene;
    scene.materials.push_back(Material{RB_MAT_PHONG, {0.7f, 0.7f, 0.6f}, {0.1f, 0.1f, 0.1f}, {}, 20.0f, 1.0f});
    scene.materials.push_back(Material{RB_MAT_PHONG, {0.8f, 0.8f, 0.8f}, {}, {100.0f, 80.9f, 29.8f}, 10.0f, 1.0f});
    Surface room, lamp;
    room.material = 0, lamp.material = 1;
    quad(room, {-3, -3, 0}, {6, 0, 0}, {0, 6, 0}, {0, 0, 1});   // floor
    quad(room, {-3, -3, 3}, {6, 0, 0}, {0, 6, 0}, {0, 0, -1});  // ceiling
    quad(room, {-3, 3, 0}, {6, 0, 0}, {0, 0, 3}, {0, -1, 0});   // back wall
    quad(room, {-1, -1, 0.8f}, {2, 0, 0}, {0, 2, 0}, {0, 0, 1});  // table top (casts a shadow)
    quad(lamp, {-0.3f, -0.3f, 2.9f}, {0.6f, 0, 0}, {0, 0.6f, 0}, {0, 0, -1});
    scene.surfaces = {room, lamp};

    Renderer raytracer(W, H);
    raytracer.LoadScene(scene);
    raytracer.params.M_Area = 32;
    raytracer.params.doVisibilityPass = raytracer.params.doTemporalReuse = raytracer.params.doSpatialReuse = true;
    std::vector<float> frame_data((size_t)W * H * 3), display_data((size_t)W * H * 4);
    Camera camera(W, H, 55.0f, {1.877986f, -7.724095f, 1.602229f}, {0, 0, 0});  // P/tutorials.cpp:35
    for (uint32_t frameCtr = 0; frameCtr < 4; ++frameCtr) {
      camera.setPosition({1.877986f + 0.05f * frameCtr, -7.724095f, 1.602229f});
      raytracer.produceRestir(camera, frameCtr, frame_data.data());
      double mean = 0;
      for (float v : frame_data) mean += v;
      mean /= (double)frame_data.size();
      std::printf("frame %u: image mean %.6f, total %.3f ms (gbuffer %.3f initial %.3f temporal %.3f spatial %.3f shade %.3f)\n",
                  frameCtr, mean, raytracer.timings.ms_total, raytracer.timings.ms_gbuffer, raytracer.timings.ms_initial,
                  raytracer.timings.ms_temporal, raytracer.timings.ms_spatial, raytracer.timings.ms_shade);
      if (!(mean > 0.0) || mean != mean) {
        std::printf("FAIL: empty or NaN image\n");
        return 1;
      }
      // the rest of the Producer loop: accumulate, tonemap + gamma into display_data, accumulator statistics
      raytracer.accumulateAndDisplay(frameCtr, /*tonemap*/ true, /*gammaCorrect*/ true, display_data.data());
      std::printf("         accumulatorMean %.6f accumulatorVariance %.6f display[0] = (%.3f %.3f %.3f %.1f)\n", raytracer.accumulatorMean,
                  raytracer.accumulatorVariance, display_data[0], display_data[1], display_data[2], display_data[3]);
      if (display_data[3] != 1.0f || !(raytracer.accumulatorMean > 0.0)) {
        std::printf("FAIL: display / statistics\n");
        return 1;
      }
    }
    std::printf("OK\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 2;
  }
}
