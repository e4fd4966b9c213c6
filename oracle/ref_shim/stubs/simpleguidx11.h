// Stand-in for P/simpleguidx11.h: only the static state the ReSTIR passes touch
// (P/simpleguidx11.h:27-56,64-65; definitions P/simpleguidx11.cpp:20-31). The Win32 / D3D11 / OIDN
// members of the real class are not declared.
#pragma once
#include "GBufferElement.h"
#include "Reservoir.h"
class SimpleGuiDX11 {
public:
	static Reservoir& getReservoirRead(const glm::vec<2, int>& coords) { return reservoirsPingPong[readReservoirBufferIndex][coords.y * width_ + coords.x]; }
	static Reservoir& getReservoirWrite(const glm::vec<2, int>& coords) { return reservoirsPingPong[writeReservoirBufferIndex][coords.y * width_ + coords.x]; }
	static Reservoir& getReservoirLastFrame(const glm::vec<2, int>& coords) { return reservoirsLastFrame[coords.y * width_ + coords.x]; }
	static Reservoir* reservoirsPingPong[2];
	static Reservoir* reservoirsLastFrame;
	static GBuffer gBuffer;
	static GBuffer gBufferLastFrame;
	static int readReservoirBufferIndex;
	static int writeReservoirBufferIndex;
	static int width_;
	static int height_;
	static glm::vec<2, int> debugPixel;
};
