// npe_harness.cpp — runs the reference's OWN benchmark program, src/num_particles_eval.cpp, compiled UNMODIFIED
// (its main() renamed on the compiler command line, -Dmain=ref_num_particles_eval_main).
// TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing in the product links this file.
//
// The program reads an .mcl snapshot, reduces the scan, builds the map with createTSDFMap, constructs a
// TSDFEvaluator and times evaluator.evaluate(...) for growing particle counts. Three builds of it (oracle/Makefile):
//   num_particles_eval_b200     linked against the product's drop-in CudaEvaluator (shim + libtsdfloc.so): use_cuda=true runs on
//                               the B200 through the boundary a maintainer would bind — the reference's benchmark, unchanged
//   num_particles_eval_refcuda  linked against the reference's own CUDA evaluator (src/cuda/*.cu, unmodified, sm_100a)
//   num_particles_eval_cpu      CudaEvaluator stubbed out (this file): use_cuda=false only, runs anywhere
// What is NOT the reference's: ROS (ref_stubs/ros/ros.h: parameters come from ROSPARAM_<name> environment variables) and
// libhdf5 (ref_stubs/highfive: an in-memory file). This driver fills that in-memory file from a raw chunk dump
//   int32 n_chunks | n_chunks x (cx, cy, cz) int32 | n_chunks x 64^3 uint32 TSDFValue words
// registered under the name given as <map-file>, then calls the reference's main with the same argv.
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include <highfive/H5File.hpp>
#include <tsdf_localization/cuda/cuda_evaluator.h>

int ref_num_particles_eval_main(int argc, char** argv);

#ifdef TSDF_NPE_NO_CUDA
namespace tsdf_localization
{
// TSDFEvaluator's ctor constructs a CudaEvaluator unconditionally (evaluation/tsdf_evaluator.h:78): a do-nothing one for the
// CPU-only build; asking it to evaluate is an error, like the reference's own non-CUDA Evaluator (cuda_evaluator.h:97-106).
CudaEvaluator::CudaEvaluator(CudaSubVoxelMap<FLOAT_T, FLOAT_T>& map, bool per_point, FLOAT_T a_hit, FLOAT_T a_range, FLOAT_T a_max, FLOAT_T max_range)
{
}
CudaEvaluator::~CudaEvaluator() {}
geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>&, const sensor_msgs::PointCloud2&, FLOAT_T[16])
{
  throw std::runtime_error("CUDA acceleration is not supported. Please install CUDA!");
}
geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>&, const std::vector<CudaPoint>&, FLOAT_T[16])
{
  throw std::runtime_error("CUDA acceleration is not supported. Please install CUDA!");
}
}  // namespace tsdf_localization
#endif

static bool load_chunk_dump(const std::string& path)
{
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  int32_t n = 0;
  bool ok = std::fread(&n, sizeof(n), 1, f) == 1 && n > 0 && n < (1 << 16);
  std::vector<int32_t> pos;
  if (ok)
  {
    pos.resize(3 * static_cast<size_t>(n));
    ok = std::fread(pos.data(), sizeof(int32_t), pos.size(), f) == pos.size();
  }
  if (ok)
  {
    const size_t words = 64u * 64u * 64u;
    auto& group = HighFive::stub_files()[path]["/map"];
    for (int32_t c = 0; ok && c < n; ++c)
    {
      const std::string tag = std::to_string(pos[3 * c]) + "_" + std::to_string(pos[3 * c + 1]) + "_" + std::to_string(pos[3 * c + 2]);
      std::vector<uint32_t>& d = group[tag];
      d.resize(words);
      ok = std::fread(d.data(), sizeof(uint32_t), words, f) == words;
    }
  }
  std::fclose(f);
  return ok;
}

int main(int argc, char** argv)
{
  if (argc == 3 && !load_chunk_dump(argv[2]))
  {
    std::cerr << "cannot read the chunk dump \"" << argv[2] << "\"" << std::endl;
    return 2;
  }
  try
  {
    return ref_num_particles_eval_main(argc, argv);
  }
  catch (const std::exception& ex)
  {
    std::cerr << "num_particles_eval: " << ex.what() << std::endl;
    return 1;
  }
}
