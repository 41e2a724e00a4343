// Empty stand-in: the reference's include/Astar-3D/map.h includes "tinyxml2.h" (absent from its tree) without using
// anything from it. Only on the include path of the oracle/_ref build of the reference's A* (oracle/Makefile).
