// slam_standins.cpp -- the static members of the stand-in Frame (TEST INFRASTRUCTURE, see slam_standins.hpp)
namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}  // namespace ORB_SLAM2
