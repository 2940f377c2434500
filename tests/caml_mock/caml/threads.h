#ifndef CAML_MOCK_THREADS_H
#define CAML_MOCK_THREADS_H
void caml_release_runtime_system(void);
void caml_acquire_runtime_system(void);
#endif
