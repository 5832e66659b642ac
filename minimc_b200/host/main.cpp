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
    // one process per GPU: MMC_WORLD_SIZE / MMC_RANK / MMC_DEVICE / MMC_COMM_ID_FILE (see Driver::InitCommFromEnvironment)
    driver->InitCommFromEnvironment();
    auto output_filepath = input_filepath;
    output_filepath.replace_extension(".out");
    std::cout << "Transporting " << driver->batchsize << " histories on " << driver->world_size << " GPU(s)..." << std::endl;
    const auto result = driver->Solve();
    if (const auto* k = dynamic_cast<const minimc::KEigenvalue*>(driver.get())) {
      std::cout.precision(17);
      std::cout << "k-effective = " << k->result().k_mean << " +/- " << k->result().k_std << std::endl;
      std::cout << "k-effective (collision estimator) = " << k->result().k_collision_mean << " +/- "
                << k->result().k_collision_std << std::endl;
      for (size_t c = 0; c < k->result().k_cycle.size(); c++)
        std::cout << "cycle " << c << " k " << k->result().k_cycle[c] << " bank " << k->result().bank_sizes[c]
                  << " k_collision " << k->result().k_collision_cycle[c] << std::endl;
    }
    if (driver->rank == 0) {  // every rank holds the batch's EstimatorSet after the all-reduce
      std::ofstream output_file{output_filepath};
      output_file << driver->batchsize << std::endl;
      output_file << result.to_string();
      output_file.close();
      std::cout << "Output written to " << std::filesystem::absolute(output_filepath) << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
