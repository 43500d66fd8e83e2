// oracle/_ref/sarw_ref: the reference's own main() (sarw.cpp:13, renamed by -Dmain=...) as a CLI binary.
// TEST INFRASTRUCTURE ONLY (format goldens, whole-program CPU timing).
#undef main
int sarw_reference_main(int argc, char** argv);
int main(int argc, char** argv) { return sarw_reference_main(argc, argv); }
