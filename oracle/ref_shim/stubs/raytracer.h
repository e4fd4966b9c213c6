// the GUI application class is not part of the hot path; MaterialPhong.cpp / MaterialLambert.cpp include it for nothing
#pragma once
