# The USE_CUDA block for portableRT's own CMakeLists.txt, written like its other backend blocks
# (CMakeLists.txt:22-67: option -> sources/definitions/links on the `portableRT` target).
# No enable_language(CUDA) is needed there: the device code is already compiled into libprt_b200.so
# (nvcc -gencode arch=compute_100a,code=sm_100a); intersect_cuda.hpp is plain C++17.
option(USE_CUDA "Enable the B200-native CUDA backend (libprt_b200)" OFF)
if(USE_CUDA)
  find_library(PRT_B200_LIB prt_b200 REQUIRED HINTS ${PRT_B200_ROOT}/lib ${PRT_B200_ROOT}/portablert_b200 ${PRT_B200_ROOT})
  find_path(PRT_B200_INC prt_b200.h REQUIRED HINTS ${PRT_B200_ROOT}/include)
  target_compile_definitions(portableRT PUBLIC USE_CUDA)
  target_include_directories(portableRT PUBLIC ${PRT_B200_INC})
  target_link_libraries(portableRT PUBLIC ${PRT_B200_LIB})
endif()
