/* ref_main_shim.cpp — TEST INFRASTRUCTURE (only tests/ may run the program built from this).
 *
 * The REFERENCE'S OWN kmers-build-signatures: /root/reference/src/kmers-build-signatures.cc with everything it
 * includes (signature_build.{h,tcc}, function_map.h, call_functions.{h,tcc}, kept_kmer_db.h, path_utils.h, ...),
 * unmodified and in place, over the stand-in third-party headers of oracle/refshim/ — its main(), option parsing,
 * output files (function.index, otu.index, genomes, final.kmers, distinct_functions, recall.report.d/*) and stdout.
 * The perfect-hash and NuDB outputs are stubbed out (cmph / NuDB are not in this image) and must not be requested.
 *
 * Built by `make -C oracle ref` into oracle/_ref/ref-kmers-build-signatures when /root/reference is present.
 */
#include "kmers-build-signatures.cc"
