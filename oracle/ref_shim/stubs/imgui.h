// Stand-in for Dear ImGui: the GUI callbacks of the reference compile to no-ops.
#pragma once
#define IM_ARRAYSIZE(a) ((int)(sizeof(a) / sizeof(*(a))))
struct ImVec2 { float x, y; ImVec2(float a = 0, float b = 0) : x(a), y(b) {} };
struct ImVec4 { float x, y, z, w; ImVec4(float a = 0, float b = 0, float c = 0, float d = 0) : x(a), y(b), z(c), w(d) {} };
namespace ImGui {
template <class... A> bool DragInt(A&&...) { return false; }
template <class... A> bool DragInt2(A&&...) { return false; }
template <class... A> bool DragFloat(A&&...) { return false; }
template <class... A> bool DragFloat2(A&&...) { return false; }
template <class... A> bool DragFloat3(A&&...) { return false; }
template <class... A> bool ColorEdit3(A&&...) { return false; }
template <class... A> bool Checkbox(A&&...) { return false; }
template <class... A> bool Combo(A&&...) { return false; }
template <class... A> bool CollapsingHeader(A&&...) { return false; }
template <class... A> bool Button(A&&...) { return false; }
template <class... A> bool InputText(A&&...) { return false; }
template <class... A> void Text(A&&...) {}
inline void Spacing() {}
inline void Separator() {}
inline void SameLine() {}
}
