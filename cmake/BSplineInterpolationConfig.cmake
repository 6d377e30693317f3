# Package configuration that lets a project written against the reference,
#
#     find_package(BSplineInterpolation)
#     target_link_libraries(main BSplineInterpolation)        (reference README.md:24-33)
#
# build against the B200 implementation unchanged: point CMake at this directory
# (-DBSplineInterpolation_DIR=<repo>/cmake).  The target carries the include directory with the
# <BSplineInterpolation/...> forwarders, the C++17 requirement of the drop-in headers and the
# device library.  Build the library first: python -m bsplineinterpolation_b200.build
get_filename_component(_bspl_b200_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(_bspl_b200_lib "${_bspl_b200_root}/bsplineinterpolation_b200/libbspline_b200.so")
if(NOT EXISTS "${_bspl_b200_lib}")
    set(BSplineInterpolation_FOUND FALSE)
    set(BSplineInterpolation_NOT_FOUND_MESSAGE
        "libbspline_b200.so has not been built (python -m bsplineinterpolation_b200.build)")
    return()
endif()
if(NOT TARGET BSplineInterpolation)
    add_library(BSplineInterpolation INTERFACE IMPORTED)
    set_target_properties(BSplineInterpolation PROPERTIES
        INTERFACE_INCLUDE_DIRECTORIES "${_bspl_b200_root}/include"
        INTERFACE_LINK_LIBRARIES "${_bspl_b200_lib}"
        INTERFACE_COMPILE_FEATURES cxx_std_17)
endif()
set(BSplineInterpolation_FOUND TRUE)
