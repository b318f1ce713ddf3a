/* Stand-in for <cmph.h> (see README.md): declarations only, so that src/perfect_hash.h and src/cmph_kmer.h
 * compile.  The perfect-hash output of the reference command line is never requested by the tests; every
 * function aborts if it is called after all. */
#pragma once
#include <cstdio>
#include <cstdlib>
typedef struct cmph_t cmph_t;
typedef struct cmph_config_t cmph_config_t;
typedef struct cmph_io_adapter_t cmph_io_adapter_t;
enum { CMPH_BDZ = 5 };
inline cmph_io_adapter_t *cmph_io_vector_adapter(char **, unsigned) { std::abort(); }
inline cmph_config_t *cmph_config_new(cmph_io_adapter_t *) { std::abort(); }
inline void cmph_config_set_algo(cmph_config_t *, int) { std::abort(); }
inline void cmph_config_set_mphf_fd(cmph_config_t *, FILE *) { std::abort(); }
inline cmph_t *cmph_new(cmph_config_t *) { std::abort(); }
inline unsigned cmph_size(cmph_t *) { std::abort(); }
inline unsigned cmph_search(cmph_t *, const char *, unsigned) { std::abort(); }
inline void cmph_config_destroy(cmph_config_t *) { std::abort(); }
inline int cmph_dump(cmph_t *, FILE *) { std::abort(); }
inline void cmph_destroy(cmph_t *) { std::abort(); }
inline cmph_t *cmph_load(FILE *) { std::abort(); }
