// Test infrastructure: compiles the REFERENCE's neighbourhood tool from where it lies
// (REF_NEIB = path of barcode_analysis/5_steps_neibourhoods/neibourhoods.cpp, set by the Makefile)
// behind a command line, without copying or editing it: its own main() is renamed away and
// read_do_and_write (neibourhoods.cpp:58-103) is called with the arguments given here.
//   usage: neibourhoods_ref <input.txt> <output.txt> <radius> <classic 0|1>
#define main ref_main_not_used
#include REF_NEIB
#undef main
#include <cstdlib>
int main(int argc, char** argv) {
    if (argc < 5) return 2;
    return read_do_and_write(argv[1], argv[2], (size_t)std::atoi(argv[3]), std::atoi(argv[4]) != 0);
}
