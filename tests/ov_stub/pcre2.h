// pcre2.h — hand-declared prototypes of the PCRE2 8-bit C API (TEST INFRASTRUCTURE ONLY).
// The image has the runtime library libpcre2-8.so.0 (10.42) but no development header; the reference pins PCRE2 10.46
// (src/CMakeLists.txt:187) and compiles with PCRE2_CODE_UNIT_WIDTH=8.  Only the entry points and option bits the reference
// uses (src/utils.cpp:256-461) are declared; the values are the library's stable ABI constants.
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef uint8_t PCRE2_UCHAR8;
typedef const PCRE2_UCHAR8* PCRE2_SPTR8;
typedef size_t PCRE2_SIZE;
typedef struct pcre2_real_code_8 pcre2_code_8;
typedef struct pcre2_real_match_data_8 pcre2_match_data_8;
typedef struct pcre2_real_compile_context_8 pcre2_compile_context_8;
typedef struct pcre2_real_general_context_8 pcre2_general_context_8;
typedef struct pcre2_real_match_context_8 pcre2_match_context_8;

#define PCRE2_UCHAR PCRE2_UCHAR8
#define PCRE2_SPTR PCRE2_SPTR8
#define pcre2_code pcre2_code_8
#define pcre2_match_data pcre2_match_data_8

#define PCRE2_UTF 0x00080000u
#define PCRE2_UCP 0x00020000u
#define PCRE2_NO_UTF_CHECK 0x40000000u
#define PCRE2_JIT_COMPLETE 0x00000001u
#define PCRE2_SUBSTITUTE_GLOBAL 0x00000100u
#define PCRE2_ERROR_NOMATCH (-1)
#define PCRE2_ERROR_NOMEMORY (-48)
#define PCRE2_ZERO_TERMINATED (~(PCRE2_SIZE)0)

pcre2_code_8* pcre2_compile_8(PCRE2_SPTR8, PCRE2_SIZE, uint32_t, int*, PCRE2_SIZE*, pcre2_compile_context_8*);
void pcre2_code_free_8(pcre2_code_8*);
int pcre2_jit_compile_8(pcre2_code_8*, uint32_t);
int pcre2_get_error_message_8(int, PCRE2_UCHAR8*, PCRE2_SIZE);
pcre2_match_data_8* pcre2_match_data_create_from_pattern_8(const pcre2_code_8*, pcre2_general_context_8*);
void pcre2_match_data_free_8(pcre2_match_data_8*);
int pcre2_match_8(const pcre2_code_8*, PCRE2_SPTR8, PCRE2_SIZE, PCRE2_SIZE, uint32_t, pcre2_match_data_8*, pcre2_match_context_8*);
int pcre2_jit_match_8(const pcre2_code_8*, PCRE2_SPTR8, PCRE2_SIZE, PCRE2_SIZE, uint32_t, pcre2_match_data_8*, pcre2_match_context_8*);
PCRE2_SIZE* pcre2_get_ovector_pointer_8(pcre2_match_data_8*);
uint32_t pcre2_get_ovector_count_8(pcre2_match_data_8*);
int pcre2_substitute_8(const pcre2_code_8*, PCRE2_SPTR8, PCRE2_SIZE, PCRE2_SIZE, uint32_t, pcre2_match_data_8*, pcre2_match_context_8*,
                       PCRE2_SPTR8, PCRE2_SIZE, PCRE2_UCHAR8*, PCRE2_SIZE*);

#define pcre2_compile pcre2_compile_8
#define pcre2_code_free pcre2_code_free_8
#define pcre2_jit_compile pcre2_jit_compile_8
#define pcre2_get_error_message pcre2_get_error_message_8
#define pcre2_match_data_create_from_pattern pcre2_match_data_create_from_pattern_8
#define pcre2_match_data_free pcre2_match_data_free_8
#define pcre2_match pcre2_match_8
#define pcre2_jit_match pcre2_jit_match_8
#define pcre2_get_ovector_pointer pcre2_get_ovector_pointer_8
#define pcre2_get_ovector_count pcre2_get_ovector_count_8
#define pcre2_substitute pcre2_substitute_8
#ifdef __cplusplus
}
#endif
