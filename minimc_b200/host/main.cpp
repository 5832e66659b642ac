// runminimc_b200 <input.xml>: the reference's CLI (minimc.cpp:10-25) on the GPU
// path.  Writes <input>.out = "<batchsize>\n" + EstimatorSet::to_string().
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "minimc.hpp"

int main(int argc, char* argv[]) {
  std::cout << "Welcome to MiniMC!" << std::endl;
  if (argc != 2) {
    std::cerr << "MiniMC accepts exactly one argument" << std::endl;
    return 2;
  }
  try {
    const std::filesystem::path input_filepath{argv[1]};
    auto driver = minimc::Driver::Create(input_filepath.string());
    auto output_filepath = input_filepath;
    std::ofstream output_file{output_filepath.replace_extension(".out")};
    std::cout << "Transporting " << driver->batchsize << " histories on the GPU..." << std::endl;
    const auto result = driver->Solve();
    output_file << driver->batchsize << std::endl;
    output_file << result.to_string();
    output_file.close();
    if (const auto* k = dynamic_cast<const minimc::KEigenvalue*>(driver.get()))
      std::cout << "k-effective = " << k->result().k_mean << " +/- " << k->result().k_std << std::endl;
    std::cout << "Output written to " << std::filesystem::absolute(output_filepath) << std::endl;
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
