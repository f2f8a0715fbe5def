// stand-in: ORBmatcher.cc includes this header but uses nothing from it
#pragma once
