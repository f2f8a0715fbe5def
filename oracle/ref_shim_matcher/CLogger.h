// stand-in for code/include/CLogger.h (spdlog): ORBmatcher.cc includes it but logs nothing
#pragma once
